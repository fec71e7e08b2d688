/*
 * gb_oracle.c -- CPU restatement of Grok's tile-coding hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (grokimagecompression_b200/, include/) may
 * link, import or call this file; it is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py as the CHECKER of the CUDA kernels.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the compiled,
 * unmodified reference (oracle/_ref/libgrok_ref.so through oracle/ref_driver.cpp) by
 * tests/test_oracle_vs_reference.py, and against the golden vectors in tests/golden/ that
 * tests/golden/make_golden.py generated from the reference.
 *
 * Written from ISO/IEC 15444-1 Annex C (MQ coder), D (coefficient bit modelling), F (DWT),
 * G (level shift / MCT) in plain scalar C, one sample at a time, with the reference's
 * deviations from the standard (fixed-point 9/7 analysis, 6 fractional magnitude bits, rate
 * bookkeeping) reproduced where the reference defines the bytes.  Each function cites the
 * reference lines it mirrors (paths relative to /root/reference/src/lib/jp2).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define GBO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Level shift + multi-component transforms                                                   */
/* ------------------------------------------------------------------------------------------ */

/* (int64(a)*b + 2^12) >> 13 : util/grok_intmath.h:209-221 */
static inline int32_t fix13(int32_t a, int32_t b) {
	return (int32_t) (((int64_t) a * (int64_t) b + 4096) >> 13);
}

/* TileProcessor.cpp:1449-1471 : subtract the DC offset; the 9/7 path also scales by 2^11 */
GBO_API void gbo_dc_shift_fwd(int32_t *x, uint64_t n, int32_t shift, int reversible) {
	for (uint64_t i = 0; i < n; ++i)
		x[i] = reversible ? x[i] - shift : (x[i] - shift) * 2048;
}

/* mct.cpp:125-135 (RCT) */
GBO_API void gbo_rct_fwd(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) {
	for (uint64_t i = 0; i < n; ++i) {
		int32_t r = c0[i], g = c1[i], b = c2[i];
		c0[i] = (r + 2 * g + b) >> 2;
		c1[i] = b - g;
		c2[i] = r - g;
	}
}

/* mct.cpp:180-190 */
GBO_API void gbo_rct_inv(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) {
	for (uint64_t i = 0; i < n; ++i) {
		int32_t y = c0[i], u = c1[i], v = c2[i];
		int32_t g = y - ((u + v) >> 2);
		c0[i] = v + g;
		c1[i] = g;
		c2[i] = u + g;
	}
}

/* mct.cpp:336-346 : ICT analysis in 13-bit fixed point */
GBO_API void gbo_ict_fwd(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) {
	for (uint64_t i = 0; i < n; ++i) {
		int32_t r = c0[i], g = c1[i], b = c2[i];
		c0[i] = fix13(r, 2449) + fix13(g, 4809) + fix13(b, 934);
		c1[i] = -fix13(r, 1382) - fix13(g, 2714) + fix13(b, 4096);
		c2[i] = fix13(r, 4096) - fix13(g, 3430) - fix13(b, 666);
	}
}

/* mct.cpp:394-404 : ICT synthesis in float, multiply then add (no fused multiply-add) */
GBO_API void gbo_ict_inv(float *c0, float *c1, float *c2, uint64_t n) {
	for (uint64_t i = 0; i < n; ++i) {
		volatile float y = c0[i], u = c1[i], v = c2[i];
		volatile float t0 = v * 1.402f;
		volatile float t1 = u * 0.34413f;
		volatile float t2 = v * 0.71414f;
		volatile float t3 = u * 1.772f;
		volatile float g0 = y - t1;
		c0[i] = y + t0;
		c1[i] = g0 - t2;
		c2[i] = y + t3;
	}
}

/* TileProcessor.cpp:1377-1432 : inverse level shift and clamp; 9/7 rounds half to even first */
GBO_API void gbo_dc_shift_inv(int32_t *x, uint64_t n, int32_t shift, int reversible, int32_t lo, int32_t hi) {
	for (uint64_t i = 0; i < n; ++i) {
		int32_t v;
		if (reversible)
			v = x[i];
		else {
			float f;
			memcpy(&f, &x[i], 4);
			v = (int32_t) lrintf(f);
		}
		v += shift;
		x[i] = v < lo ? lo : (v > hi ? hi : v);
	}
}

/* ------------------------------------------------------------------------------------------ */
/* Wavelets.  One line = samples at canvas coordinates c0 .. c0+len-1; even coordinates are   */
/* low-pass, odd are high-pass; whole-sample symmetric extension at both ends.                */
/* ------------------------------------------------------------------------------------------ */

static inline int mirror(int i, int len) {
	/* reflect an index that is at most one step outside [0,len) */
	if (i < 0) return -i;
	if (i >= len) return 2 * (len - 1) - i;
	return i;
}

/* dwt53.cpp:150-169 : forward 5/3 on an interleaved line (x[0] is at parity `cas`) */
static void fwd53_line(int32_t *x, int len, int cas) {
	if (len == 1) {
		if (cas) x[0] *= 2; /* lone high-pass sample */
		return;
	}
	for (int i = 1 - cas; i < len; i += 2) /* predict: high-pass at odd coordinates */
		x[i] -= (x[mirror(i - 1, len)] + x[mirror(i + 1, len)]) >> 1;
	for (int i = cas; i < len; i += 2) /* update: low-pass at even coordinates */
		x[i] += (x[mirror(i - 1, len)] + x[mirror(i + 1, len)] + 2) >> 2;
}

/* dwt.cpp:256-363 : inverse 5/3 */
static void inv53_line(int32_t *x, int len, int cas) {
	if (len == 1) {
		if (cas) x[0] /= 2; /* C division, truncates toward zero (dwt.cpp:349) */
		return;
	}
	for (int i = cas; i < len; i += 2)
		x[i] -= (x[mirror(i - 1, len)] + x[mirror(i + 1, len)] + 2) >> 2;
	for (int i = 1 - cas; i < len; i += 2)
		x[i] += (x[mirror(i - 1, len)] + x[mirror(i + 1, len)]) >> 1;
}

/* dwt97.cpp:90-123 : forward 9/7 in 13-bit fixed point */
static void fwd97_line(int32_t *x, int len, int cas) {
	if (len == 1)
		return;
	for (int i = 1 - cas; i < len; i += 2)
		x[i] -= fix13(x[mirror(i - 1, len)] + x[mirror(i + 1, len)], 12994);
	for (int i = cas; i < len; i += 2)
		x[i] -= fix13(x[mirror(i - 1, len)] + x[mirror(i + 1, len)], 434);
	for (int i = 1 - cas; i < len; i += 2)
		x[i] += fix13(x[mirror(i - 1, len)] + x[mirror(i + 1, len)], 7233);
	for (int i = cas; i < len; i += 2)
		x[i] += fix13(x[mirror(i - 1, len)] + x[mirror(i + 1, len)], 3633);
	for (int i = 1 - cas; i < len; i += 2)
		x[i] = fix13(x[i], 5039);
	for (int i = cas; i < len; i += 2)
		x[i] = fix13(x[i], 6659);
}

/* dwt.cpp:172-178, 1413-1537 : inverse 9/7 in float; every step is  x + (l + r) * c  with a
 * rounded multiply followed by a rounded add */
static void inv97_step(float *x, int len, int first, float c) {
	for (int i = first; i < len; i += 2) {
		volatile float s = x[mirror(i - 1, len)] + x[mirror(i + 1, len)];
		volatile float p = s * c;
		x[i] = x[i] + p;
	}
}
static void inv97_line(float *x, int len, int cas) {
	if (len == 1)
		return;
	for (int i = cas; i < len; i += 2) {
		volatile float p = x[i] * 1.230174105f;
		x[i] = p;
	}
	for (int i = 1 - cas; i < len; i += 2) {
		volatile float p = x[i] * 1.625732422f;
		x[i] = p;
	}
	inv97_step(x, len, cas, -0.443506852f);
	inv97_step(x, len, 1 - cas, -0.882911075f);
	inv97_step(x, len, cas, 0.052980118f);
	inv97_step(x, len, 1 - cas, 1.586134342f);
}

static inline uint32_t cdiv2n(uint32_t a, uint32_t n) {
	return (uint32_t) (((uint64_t) a + ((1ull << n) - 1)) >> n);
}

/* WaveletForward.h:40-161 : per level, finest first: every column, then every row; low half
 * first (Mallat layout) in place, row stride = tile-component width. */
GBO_API int gbo_dwt_fwd(int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
		int reversible) {
	uint32_t stride = x1 - x0;
	uint32_t maxlen = (x1 - x0) > (y1 - y0) ? (x1 - x0) : (y1 - y0);
	int32_t *line = (int32_t*) malloc(sizeof(int32_t) * (maxlen + 2));
	int32_t *tmp = (int32_t*) malloc(sizeof(int32_t) * (maxlen + 2));
	if (!line || !tmp)
		return 1;
	for (uint32_t lvl = 0; lvl + 1 < numres; ++lvl) {
		uint32_t rx0 = cdiv2n(x0, lvl), rx1 = cdiv2n(x1, lvl);
		uint32_t ry0 = cdiv2n(y0, lvl), ry1 = cdiv2n(y1, lvl);
		uint32_t rw = rx1 - rx0, rh = ry1 - ry0;
		uint32_t sw = cdiv2n(x1, lvl + 1) - cdiv2n(x0, lvl + 1); /* low-pass count horizontally */
		uint32_t sh = cdiv2n(y1, lvl + 1) - cdiv2n(y0, lvl + 1);
		int cas_c = (int) (ry0 & 1), cas_r = (int) (rx0 & 1);
		for (uint32_t c = 0; c < rw && rh; ++c) {
			for (uint32_t k = 0; k < rh; ++k)
				line[k] = buf[(size_t) k * stride + c];
			if (reversible) fwd53_line(line, (int) rh, cas_c); else fwd97_line(line, (int) rh, cas_c);
			for (uint32_t k = 0; k < rh; ++k) { /* de-interleave: dwt_utils.cpp:84-105 */
				uint32_t even = ((k & 1) == (uint32_t) cas_c);
				uint32_t dst = even ? (k >> 1) : sh + (k >> 1);
				buf[(size_t) dst * stride + c] = line[k];
			}
		}
		for (uint32_t r = 0; r < rh && rw; ++r) {
			int32_t *row = buf + (size_t) r * stride;
			memcpy(line, row, sizeof(int32_t) * rw);
			if (reversible) fwd53_line(line, (int) rw, cas_r); else fwd97_line(line, (int) rw, cas_r);
			for (uint32_t k = 0; k < rw; ++k) {
				uint32_t even = ((k & 1) == (uint32_t) cas_r);
				uint32_t dst = even ? (k >> 1) : sw + (k >> 1);
				tmp[dst] = line[k];
			}
			memcpy(row, tmp, sizeof(int32_t) * rw);
		}
	}
	free(line);
	free(tmp);
	return 0;
}

/* dwt.cpp:724-858 (5/3) and 1544-1738 (9/7): per level, coarsest first: every row of the new
 * resolution, then every column.  Row stride = width of the highest DECODED resolution
 * (dwt.cpp:735-736), which is how reduced-resolution decode works. */
GBO_API int gbo_dwt_inv(int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
		uint32_t numres_decode, int reversible) {
	uint32_t top = numres - numres_decode; /* decomposition level of the output */
	uint32_t stride = cdiv2n(x1, top) - cdiv2n(x0, top);
	uint32_t h = cdiv2n(y1, top) - cdiv2n(y0, top);
	uint32_t maxlen = stride > h ? stride : h;
	int32_t *line = (int32_t*) malloc(sizeof(int32_t) * (maxlen + 2));
	if (!line)
		return 1;
	for (int lvl = (int) numres - 2; lvl >= (int) top; --lvl) {
		uint32_t rx0 = cdiv2n(x0, lvl), rx1 = cdiv2n(x1, lvl);
		uint32_t ry0 = cdiv2n(y0, lvl), ry1 = cdiv2n(y1, lvl);
		uint32_t rw = rx1 - rx0, rh = ry1 - ry0;
		uint32_t sw = cdiv2n(x1, lvl + 1) - cdiv2n(x0, lvl + 1);
		uint32_t sh = cdiv2n(y1, lvl + 1) - cdiv2n(y0, lvl + 1);
		int cas_c = (int) (ry0 & 1), cas_r = (int) (rx0 & 1);
		for (uint32_t r = 0; r < rh && rw; ++r) {
			int32_t *row = buf + (size_t) r * stride;
			for (uint32_t k = 0; k < rw; ++k) {
				uint32_t even = ((k & 1) == (uint32_t) cas_r);
				line[k] = row[even ? (k >> 1) : sw + (k >> 1)];
			}
			if (reversible) inv53_line(line, (int) rw, cas_r); else inv97_line((float*) line, (int) rw, cas_r);
			memcpy(row, line, sizeof(int32_t) * rw);
		}
		for (uint32_t c = 0; c < rw && rh; ++c) {
			for (uint32_t k = 0; k < rh; ++k) {
				uint32_t even = ((k & 1) == (uint32_t) cas_c);
				line[k] = buf[(size_t) (even ? (k >> 1) : sh + (k >> 1)) * stride + c];
			}
			if (reversible) inv53_line(line, (int) rh, cas_c); else inv97_line((float*) line, (int) rh, cas_c);
			for (uint32_t k = 0; k < rh; ++k)
				buf[(size_t) k * stride + c] = line[k];
		}
	}
	free(line);
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Quantisation (T1Part1.cpp:45-95) and de-quantisation (T1Part1.cpp:216-329)                 */
/* ------------------------------------------------------------------------------------------ */

/* Copies a w x h block out of the tile plane into `out` with 6 fractional bits; returns max|q|.
 * 5/3: q = x * 64.  9/7: q = (x * inv_step + 2^17) >> 18 with inv_step in 13-bit fixed point. */
GBO_API uint32_t gbo_quantise_block(const int32_t *tile, uint32_t tile_stride, uint32_t w, uint32_t h,
		int reversible, int32_t inv_step, int32_t *out) {
	uint32_t mx = 0;
	for (uint32_t j = 0; j < h; ++j)
		for (uint32_t i = 0; i < w; ++i) {
			int32_t x = tile[(size_t) j * tile_stride + i];
			int32_t q = reversible ? x * 64 : (int32_t) (((int64_t) x * inv_step + (1 << 17)) >> 18);
			uint32_t a = (uint32_t) (q < 0 ? -q : q);
			if (a > mx) mx = a;
			out[j * w + i] = q;
		}
	return mx;
}

/* Scatter a decoded block (values with one extra low bit) into the tile plane: 5/3 halves with
 * C division; 9/7 multiplies by the (already halved, TileComponent.cpp:336) band step size. */
GBO_API void gbo_dequantise_block(const int32_t *blk, uint32_t w, uint32_t h, int reversible, float stepsize,
		int32_t *tile, uint32_t tile_stride) {
	for (uint32_t j = 0; j < h; ++j)
		for (uint32_t i = 0; i < w; ++i) {
			int32_t v = blk[j * w + i];
			if (reversible)
				tile[(size_t) j * tile_stride + i] = v / 2;
			else {
				volatile float f = (float) v * stepsize;
				float g = f;
				memcpy(&tile[(size_t) j * tile_stride + i], &g, 4);
			}
		}
}

/* ------------------------------------------------------------------------------------------ */
/* MQ coder, ISO 15444-1 Annex C.  Table C.2: Qe, NMPS, NLPS, SWITCH.                         */
/* (reference keeps the same table doubled by MPS value: mqc_enc.cpp:69-164)                  */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint16_t qe; uint8_t nmps, nlps, sw; } mq_row;
static const mq_row MQ[47] = {
	{0x5601, 1, 1, 1}, {0x3401, 2, 6, 0}, {0x1801, 3, 9, 0}, {0x0AC1, 4, 12, 0}, {0x0521, 5, 29, 0},
	{0x0221, 38, 33, 0}, {0x5601, 7, 6, 1}, {0x5401, 8, 14, 0}, {0x4801, 9, 14, 0}, {0x3801, 10, 14, 0},
	{0x3001, 11, 17, 0}, {0x2401, 12, 18, 0}, {0x1C01, 13, 20, 0}, {0x1601, 29, 21, 0}, {0x5601, 15, 14, 1},
	{0x5401, 16, 14, 0}, {0x5101, 17, 15, 0}, {0x4801, 18, 16, 0}, {0x3801, 19, 17, 0}, {0x3401, 20, 18, 0},
	{0x3001, 21, 19, 0}, {0x2801, 22, 19, 0}, {0x2401, 23, 20, 0}, {0x2201, 24, 21, 0}, {0x1C01, 25, 22, 0},
	{0x1801, 26, 23, 0}, {0x1601, 27, 24, 0}, {0x1401, 28, 25, 0}, {0x1201, 29, 26, 0}, {0x1101, 30, 27, 0},
	{0x0AC1, 31, 28, 0}, {0x09C1, 32, 29, 0}, {0x08A1, 33, 30, 0}, {0x0521, 34, 31, 0}, {0x0441, 35, 32, 0},
	{0x02A1, 36, 33, 0}, {0x0221, 37, 34, 0}, {0x0141, 38, 35, 0}, {0x0111, 39, 36, 0}, {0x0085, 40, 37, 0},
	{0x0049, 41, 38, 0}, {0x0025, 42, 39, 0}, {0x0015, 43, 40, 0}, {0x0009, 44, 41, 0}, {0x0005, 45, 42, 0},
	{0x0001, 45, 43, 0}, {0x5601, 46, 46, 0}
};

enum { STY_LAZY = 1, STY_RESET = 2, STY_TERMALL = 4, STY_VSC = 8, STY_PTERM = 16, STY_SEGSYM = 32 };
enum { CTX_ZC0 = 0, CTX_SC0 = 9, CTX_MR0 = 14, CTX_AGG = 17, CTX_UNI = 18, NCTX = 19 };

typedef struct {
	uint32_t a, c;
	int ct;
	uint8_t *start, *bp; /* bp starts one byte BEFORE start (mqc_enc.cpp:245-257) */
	uint8_t st[NCTX], mps[NCTX];
} mqe;

/* mqc_dec.cpp:207-214 */
static void mq_reset_ctx(uint8_t *st, uint8_t *mps) {
	memset(st, 0, NCTX);
	memset(mps, 0, NCTX);
	st[CTX_ZC0] = 4;
	st[CTX_AGG] = 3;
	st[CTX_UNI] = 46;
}

static void mqe_init(mqe *q, uint8_t *start) {
	q->a = 0x8000;
	q->c = 0;
	q->ct = 12;
	q->start = start;
	q->bp = start - 1;
	mq_reset_ctx(q->st, q->mps);
}

/* Figure C.9 BYTEOUT (mqc_enc.cpp:168-199) */
static void mqe_byteout(mqe *q) {
	if (*q->bp != 0xFF && (q->c & 0x8000000)) { /* propagate the carry into the last byte */
		(*q->bp)++;
		q->c &= 0x7FFFFFF;
	}
	if (*q->bp == 0xFF) { /* bit stuffing: next byte carries only 7 bits */
		*++q->bp = (uint8_t) (q->c >> 20);
		q->c &= 0xFFFFF;
		q->ct = 7;
	} else {
		*++q->bp = (uint8_t) (q->c >> 19);
		q->c &= 0x7FFFF;
		q->ct = 8;
	}
}

static void mqe_renorm(mqe *q) {
	do {
		q->a <<= 1;
		q->c <<= 1;
		if (--q->ct == 0)
			mqe_byteout(q);
	} while (!(q->a & 0x8000));
}

/* Figures C.6-C.8 CODEMPS / CODELPS (mqc_enc.cpp:211-233, 259-264) */
static void mqe_encode(mqe *q, int cx, int d) {
	const mq_row *r = &MQ[q->st[cx]];
	uint32_t qe = r->qe;
	q->a -= qe;
	if (d == q->mps[cx]) {
		if (q->a & 0x8000) {
			q->c += qe;
			return;
		}
		if (q->a < qe) q->a = qe; else q->c += qe;
		q->st[cx] = r->nmps;
	} else {
		if (q->a < qe) q->c += qe; else q->a = qe;
		if (r->sw) q->mps[cx] ^= 1;
		q->st[cx] = r->nlps;
	}
	mqe_renorm(q);
}

/* Figure C.11 FLUSH (mqc_enc.cpp:235-243, 274-287) */
static void mqe_flush(mqe *q) {
	uint32_t t = q->c + q->a;
	q->c |= 0xFFFF;
	if (q->c >= t) q->c -= 0x8000;
	q->c <<= q->ct;
	mqe_byteout(q);
	q->c <<= q->ct;
	mqe_byteout(q);
	if (*q->bp != 0xFF) q->bp++;
}

static inline uint32_t mqe_numbytes(const mqe *q) { return (uint32_t) (q->bp - q->start); }

/* predictable termination (mqc_enc.cpp:396-407) */
static void mqe_erterm(mqe *q) {
	int k = 11 - q->ct + 1;
	while (k > 0) {
		q->c <<= q->ct;
		q->ct = 0;
		mqe_byteout(q);
		k -= q->ct;
	}
	if (*q->bp != 0xFF) mqe_byteout(q);
}

/* INITENC again after a terminated pass: the last byte of the previous segment becomes the pending byte
 * (mqc_enc.cpp:379-394) */
static void mqe_restart(mqe *q) {
	q->a = 0x8000;
	q->c = 0;
	q->ct = 12;
	q->bp--;
	if (*q->bp == 0xFF) q->ct = 13;
}

/* selective arithmetic coding bypass: raw bits with the same bit stuffing (mqc_enc.cpp:291-377) */
#define MQE_BYPASS_CT_INIT 0x7FFFFFFF
static void mqe_bypass_init(mqe *q) {
	q->c = 0;
	q->ct = MQE_BYPASS_CT_INIT; /* "no bit written yet" */
}
static void mqe_bypass(mqe *q, int d) {
	if (q->ct == MQE_BYPASS_CT_INIT) q->ct = 8;
	q->ct--;
	q->c += (uint32_t) d << q->ct;
	if (q->ct == 0) {
		*q->bp = (uint8_t) q->c;
		q->ct = 8;
		if (*q->bp == 0xFF) q->ct = 7; /* the next byte keeps its MSB clear */
		q->bp++;
		q->c = 0;
	}
}
static int mqe_bypass_pending(const mqe *q, int erterm) {
	return q->ct < 7 || (q->ct == 7 && (erterm || q->bp[-1] != 0xFF));
}
static uint32_t mqe_bypass_extra_bytes(const mqe *q, int erterm) { return mqe_bypass_pending(q, erterm) ? 2 : 1; }
static void mqe_bypass_flush(mqe *q, int erterm) {
	if (mqe_bypass_pending(q, erterm)) {
		int bit = 0;
		while (q->ct > 0) { /* pad with 0,1,0,1,... */
			q->ct--;
			q->c += (uint32_t) bit << q->ct;
			bit = 1 - bit;
		}
		*q->bp = (uint8_t) q->c;
		q->bp++;
	} else if (q->ct == 7 && q->bp[-1] == 0xFF) {
		q->bp--; /* a trailing 0xFF is dropped */
	} else if (q->ct == 8 && !erterm && q->bp[-1] == 0x7F && q->bp[-2] == 0xFF) {
		q->bp -= 2; /* FF 7F reads the same as the end marker the decoder synthesises */
	}
}

typedef struct {
	uint32_t a, c;
	int ct;
	const uint8_t *buf;
	uint32_t len, pos;
	uint8_t st[NCTX], mps[NCTX];
} mqd;

/* bytes past the end of the segment read as 0xFF (the reference plants an FF FF marker there:
 * mqc_dec.cpp:161-177) */
static inline uint32_t mqd_byte(const mqd *q, uint32_t i) { return i < q->len ? q->buf[i] : 0xFF; }

/* mqc_dec_inl.h:114-134 BYTEIN */
static void mqd_bytein(mqd *q) {
	uint32_t next = mqd_byte(q, q->pos + 1);
	if (mqd_byte(q, q->pos) == 0xFF) {
		if (next > 0x8F) {
			q->c += 0xFF00;
			q->ct = 8;
		} else {
			q->pos++;
			q->c += next << 9;
			q->ct = 7;
		}
	} else {
		q->pos++;
		q->c += next << 8;
		q->ct = 8;
	}
}

/* mqc_dec.cpp:179-201 INITDEC */
static void mqd_init(mqd *q, const uint8_t *buf, uint32_t len) {
	q->buf = buf;
	q->len = len;
	q->pos = 0;
	q->c = mqd_byte(q, 0) << 16;
	mqd_bytein(q);
	q->c <<= 7;
	q->ct -= 7;
	q->a = 0x8000;
	mq_reset_ctx(q->st, q->mps);
}

/* mqc_dec_inl.h:60-86, 136-169 DECODE */
static int mqd_decode(mqd *q, int cx) {
	const mq_row *r = &MQ[q->st[cx]];
	uint32_t qe = r->qe;
	int d;
	q->a -= qe;
	if ((q->c >> 16) < qe) { /* LPS sub-interval selected */
		if (q->a < qe) { d = q->mps[cx]; q->st[cx] = r->nmps; }
		else { d = q->mps[cx] ^ 1; if (r->sw) q->mps[cx] ^= 1; q->st[cx] = r->nlps; }
		q->a = qe;
	} else {
		q->c -= qe << 16;
		if (q->a & 0x8000)
			return q->mps[cx];
		if (q->a < qe) { d = q->mps[cx] ^ 1; if (r->sw) q->mps[cx] ^= 1; q->st[cx] = r->nlps; }
		else { d = q->mps[cx]; q->st[cx] = r->nmps; }
	}
	do {
		if (q->ct == 0) mqd_bytein(q);
		q->a <<= 1;
		q->c <<= 1;
		q->ct--;
	} while (q->a < 0x8000);
	return d;
}

/* ------------------------------------------------------------------------------------------ */
/* Coefficient bit modelling, Annex D, on per-sample state arrays with a one-sample border.    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
	int w, h, S; /* S = w+2 : row pitch of the state arrays */
	uint8_t *sig, *neg, *vis, *refd; /* significant, negative, visited this plane, refined before */
	int orient;
	int vsc; /* vertically stripe-causal contexts: the row below a stripe counts as insignificant (t1.cpp:168-190) */
} t1s;

#define AT(p, x, y) ((p)[((y) + 1) * s->S + (x) + 1])
/* significance of a neighbour in the row below sample row y: hidden when y is the last row of its stripe and VSC is on */
#define BELOW(x, y) ((s->vsc && ((y) & 3) == 3) ? 0 : AT(s->sig, x, (y) + 1))

static int t1s_alloc(t1s *s, int w, int h, int orient) {
	s->w = w; s->h = h; s->S = w + 2; s->orient = orient;
	size_t n = (size_t) (w + 2) * (h + 2);
	s->sig = (uint8_t*) calloc(4, n);
	if (!s->sig) return 1;
	s->neg = s->sig + n; s->vis = s->neg + n; s->refd = s->vis + n;
	return 0;
}
static void t1s_free(t1s *s) { free(s->sig); }

/* Table D.1 (t1_generate_luts.cpp:63-141); the HL band swaps the roles of h and v */
static int ctx_zc(const t1s *s, int x, int y) {
	int h = AT(s->sig, x - 1, y) + AT(s->sig, x + 1, y);
	int v = AT(s->sig, x, y - 1) + BELOW(x, y);
	int d = AT(s->sig, x - 1, y - 1) + AT(s->sig, x + 1, y - 1) + BELOW(x - 1, y) + BELOW(x + 1, y);
	if (s->orient == 1) { int t = h; h = v; v = t; }
	if (s->orient == 3) {
		int hv = h + v;
		if (d == 0) return hv >= 2 ? 2 : hv;
		if (d == 1) return hv >= 2 ? 5 : 3 + hv;
		if (d == 2) return hv >= 1 ? 7 : 6;
		return 8;
	}
	if (h == 2) return 8;
	if (h == 1) return v >= 1 ? 7 : (d >= 1 ? 6 : 5);
	if (v == 2) return 4;
	if (v == 1) return 3;
	return d >= 2 ? 2 : d;
}

static int any_sig_neighbour(const t1s *s, int x, int y) {
	return AT(s->sig, x - 1, y) | AT(s->sig, x + 1, y) | AT(s->sig, x, y - 1) | BELOW(x, y)
			| AT(s->sig, x - 1, y - 1) | AT(s->sig, x + 1, y - 1) | BELOW(x - 1, y) | BELOW(x + 1, y);
}

/* Tables D.2 / D.3 (t1_generate_luts.cpp:143-209): returns context, *xorbit = sign prediction */
static int ctx_sc(const t1s *s, int x, int y, int *xorbit) {
	#define CONTRIB(xx, yy) (AT(s->sig, xx, yy) ? (AT(s->neg, xx, yy) ? -1 : 1) : 0)
	int hc = CONTRIB(x - 1, y) + CONTRIB(x + 1, y);
	int vc = CONTRIB(x, y - 1) + (BELOW(x, y) ? (AT(s->neg, x, y + 1) ? -1 : 1) : 0);
	#undef CONTRIB
	hc = hc > 1 ? 1 : (hc < -1 ? -1 : hc);
	vc = vc > 1 ? 1 : (vc < -1 ? -1 : vc);
	*xorbit = (hc < 0 || (hc == 0 && vc < 0)) ? 1 : 0;
	if (hc < 0) { hc = -hc; vc = -vc; }
	if (hc == 0) return CTX_SC0 + (vc == 0 ? 0 : 1);
	return CTX_SC0 + 3 + vc; /* vc=-1 -> 11, 0 -> 12, 1 -> 13 */
}

/* Table D.4 (t1.cpp:146-151) */
static int ctx_mr(const t1s *s, int x, int y) {
	if (AT(s->refd, x, y)) return CTX_MR0 + 2;
	return any_sig_neighbour(s, x, y) ? CTX_MR0 + 1 : CTX_MR0;
}

/* distortion-estimate tables (t1_generate_luts.cpp:291-317): index = 7 magnitude bits, the top
 * one being the bit of the current plane and the lower six the fractional bits */
static int nmsedec_sig(uint32_t mag, int bp) {
	uint32_t i = (mag >> bp) & 127;
	double t = i / 64.0, u = t, v = t - 1.5;
	double e = bp > 0 ? (u * u - v * v) : (u * u);
	int r = (int) (floor(e * 64.0 + 0.5) / 64.0 * 8192.0);
	return r < 0 ? 0 : r;
}
static int nmsedec_ref(uint32_t mag, int bp) {
	uint32_t i = (mag >> bp) & 127;
	double t = i / 64.0, u = t - 1.0, v = (i & 64) ? t - 1.5 : t - 0.5;
	double e = bp > 0 ? (u * u - v * v) : (u * u);
	int r = (int) (floor(e * 64.0 + 0.5) / 64.0 * 8192.0);
	return r < 0 ? 0 : r;
}

GBO_API void gbo_nmsedec_tables(int16_t *sig, int16_t *sig0, int16_t *ref, int16_t *ref0) {
	for (uint32_t i = 0; i < 128; ++i) {
		sig[i] = (int16_t) nmsedec_sig(i << 1, 1);
		sig0[i] = (int16_t) nmsedec_sig(i, 0);
		ref[i] = (int16_t) nmsedec_ref(i << 1, 1);
		ref0[i] = (int16_t) nmsedec_ref(i, 0);
	}
}

/* context tables in the reference's LUT index conventions, for table-level comparison:
 * zc[orient*512 + f] with f = 9 neighbourhood bits (bit k = row k/3, column k%3; bit 4 = self),
 * sc[f]/spb[f] with f = {sgnW, sigN, sgnE, sigW, sgnN, sigE, sgnS, sigS} as bits 0..7. */
GBO_API void gbo_context_tables(uint8_t *zc, uint8_t *sc, uint8_t *spb) {
	t1s st = {0}, *s = &st;
	t1s_alloc(s, 3, 3, 0);
	for (int orient = 0; orient < 4; ++orient)
		for (int f = 0; f < 512; ++f) {
			s->orient = orient;
			for (int k = 0; k < 9; ++k) AT(s->sig, k % 3, k / 3) = (uint8_t) ((f >> k) & 1);
			AT(s->sig, 1, 1) = 0;
			zc[orient * 512 + f] = (uint8_t) ctx_zc(s, 1, 1);
		}
	for (int f = 0; f < 256; ++f) {
		memset(s->sig, 0, 4 * 25);
		AT(s->sig, 0, 1) = (f >> 3) & 1; AT(s->neg, 0, 1) = f & 1;
		AT(s->sig, 1, 0) = (f >> 1) & 1; AT(s->neg, 1, 0) = (f >> 4) & 1;
		AT(s->sig, 2, 1) = (f >> 5) & 1; AT(s->neg, 2, 1) = (f >> 2) & 1;
		AT(s->sig, 1, 2) = (f >> 7) & 1; AT(s->neg, 1, 2) = (f >> 6) & 1;
		int x;
		sc[f] = (uint8_t) ctx_sc(s, 1, 1, &x);
		spb[f] = (uint8_t) x;
	}
	t1s_free(s);
}

/* ---- encoder ------------------------------------------------------------------------------ */

typedef struct {
	const int32_t *data; /* w*h, 6 fractional bits */
	t1s s;
	mqe q;
	int nmsedec;
	uint64_t nsym; /* MQ decisions (and raw bits) coded: the algorithmic work unit of Tier-1 */
	int raw;       /* this pass bypasses the arithmetic coder (cblk_sty LAZY, t1.cpp:1229-1231) */
} t1e;

static void mqe_bypass(mqe *q, int d);

/* one binary decision: MQ with its context, or a raw bit in a bypass pass (t1.cpp:208-211, 224-229, 456-460) */
static void enc_emit(t1e *e, int cx, int d) {
	if (e->raw) mqe_bypass(&e->q, d); else mqe_encode(&e->q, cx, d);
	e->nsym++;
}

static inline uint32_t mag_at(const t1e *e, int x, int y) {
	int32_t v = e->data[y * e->s.w + x];
	return (uint32_t) (v < 0 ? -v : v);
}

static void enc_sign_and_mark(t1e *e, int x, int y, int bp) {
	t1s *s = &e->s;
	int xorbit, neg = e->data[y * s->w + x] < 0;
	int cx = ctx_sc(s, x, y, &xorbit);
	e->nmsedec += nmsedec_sig(mag_at(e, x, y), bp);
	enc_emit(e, cx, e->raw ? neg : neg ^ xorbit); /* raw passes send the sign itself */
	AT(s->sig, x, y) = 1;
	AT(s->neg, x, y) = (uint8_t) neg;
}

/* t1.cpp:197-231, 287-338 */
static void enc_sigpass(t1e *e, int bp) {
	t1s *s = &e->s;
	e->nmsedec = 0;
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x)
			for (int y = y0; y < y0 + 4 && y < s->h; ++y) {
				if (AT(s->sig, x, y) || !any_sig_neighbour(s, x, y))
					continue;
				int bit = (mag_at(e, x, y) >> (bp + 6)) & 1;
				enc_emit(e, CTX_ZC0 + ctx_zc(s, x, y), bit);
				if (bit) enc_sign_and_mark(e, x, y, bp);
				AT(s->vis, x, y) = 1;
			}
}

/* t1.cpp:443-463, 498-555 */
static void enc_refpass(t1e *e, int bp) {
	t1s *s = &e->s;
	e->nmsedec = 0;
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x)
			for (int y = y0; y < y0 + 4 && y < s->h; ++y) {
				if (!AT(s->sig, x, y) || AT(s->vis, x, y))
					continue;
				e->nmsedec += nmsedec_ref(mag_at(e, x, y), bp);
				enc_emit(e, ctx_mr(s, x, y), (mag_at(e, x, y) >> (bp + 6)) & 1);
				AT(s->refd, x, y) = 1;
			}
}

/* t1.cpp:639-699, 739-782 */
static void enc_clnpass(t1e *e, int bp) {
	t1s *s = &e->s;
	e->nmsedec = 0;
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x) {
			int y = y0;
			int full = (y0 + 4 <= s->h);
			int runmode = full;
			for (int k = 0; k < 4 && runmode; ++k)
				if (AT(s->sig, x, y0 + k) || AT(s->vis, x, y0 + k) || any_sig_neighbour(s, x, y0 + k))
					runmode = 0;
			if (runmode) {
				int r = 0;
				while (r < 4 && !((mag_at(e, x, y0 + r) >> (bp + 6)) & 1)) r++;
				mqe_encode(&e->q, CTX_AGG, r != 4);
				e->nsym++;
				if (r == 4)
					continue;
				mqe_encode(&e->q, CTX_UNI, r >> 1);
				mqe_encode(&e->q, CTX_UNI, r & 1);
				e->nsym += 2;
				y = y0 + r;
				enc_sign_and_mark(e, x, y, bp); /* the 1 is implied: straight to its sign */
				y++;
			}
			for (; y < y0 + 4 && y < s->h; ++y) {
				if (AT(s->sig, x, y) || AT(s->vis, x, y))
					continue;
				int bit = (mag_at(e, x, y) >> (bp + 6)) & 1;
				mqe_encode(&e->q, CTX_ZC0 + ctx_zc(s, x, y), bit);
				e->nsym++;
				if (bit) enc_sign_and_mark(e, x, y, bp);
			}
		}
	memset(s->vis, 0, (size_t) s->S * (s->h + 2));
}

/*
 * Tier-1 encode of one block (t1.cpp:1182-1326, cblk_sty == 0).
 *  data      w*h quantised coefficients with 6 fractional bits
 *  wbase     (mct_norm * dwt_norm) * stepsize, the plane-independent factor of t1_getwmsedec
 *            (t1.cpp:912-932); ignored unless do_rd
 *  out       byte buffer, capacity >= w*h*4 + 8; out[-1] must be addressable: callers pass
 *            a pointer one past a zero pad byte
 * Returns the number of passes; rates[] are the reference's per-pass cumulative byte counts
 * (after the monotone and trailing-0xFF fix-ups), dists[] cumulative distortion decreases.
 */
GBO_API int gbo_t1_encode_block(const int32_t *data, int w, int h, int orient, int do_rd, double wbase,
		uint8_t *out, uint32_t *numbps_out, uint32_t *rates, double *dists, uint64_t *nsym_out) {
	t1e e;
	memset(&e, 0, sizeof(e));
	e.data = data;
	uint32_t mx = 0;
	for (int i = 0; i < w * h; ++i) {
		uint32_t a = (uint32_t) (data[i] < 0 ? -data[i] : data[i]);
		if (a > mx) mx = a;
	}
	int nbits = 0;
	while (mx >> nbits) nbits++; /* floorlog2(max)+1 */
	int numbps = nbits > 6 ? nbits - 6 : 0;
	*numbps_out = (uint32_t) numbps;
	if (nsym_out) *nsym_out = 0;
	if (numbps == 0)
		return 0;
	if (t1s_alloc(&e.s, w, h, orient))
		return -1;
	out[-1] = 0;
	mqe_init(&e.q, out);
	int npass = 0;
	double cum = 0.0;
	for (int bp = numbps - 1; bp >= 0; --bp)
		for (int type = (bp == numbps - 1 ? 2 : 0); type < 3; ++type) {
			if (type == 0) enc_sigpass(&e, bp);
			else if (type == 1) enc_refpass(&e, bp);
			else enc_clnpass(&e, bp);
			if (do_rd) {
				double x = wbase * (double) (1 << bp);
				x *= x * e.nmsedec / 8192.0;
				cum += x;
				dists[npass] = cum;
			} else
				dists[npass] = 0.0;
			if (type == 2 && bp == 0) {
				mqe_flush(&e.q);
				rates[npass] = mqe_numbytes(&e.q);
			} else
				rates[npass] = mqe_numbytes(&e.q) + (e.q.ct < 5 ? 6 : 5); /* t1.cpp:1278-1288 */
			npass++;
		}
	uint32_t last = mqe_numbytes(&e.q);
	for (int i = npass - 1; i >= 0; --i) { /* t1.cpp:1303-1313 */
		if (rates[i] > last) rates[i] = last; else last = rates[i];
	}
	for (int i = 0; i < npass; ++i) /* t1.cpp:1315-1324 */
		if (out[(int) rates[i] - 1] == 0xFF) rates[i]--;
	if (nsym_out) *nsym_out = e.nsym;
	t1s_free(&e.s);
	return npass;
}

/* t1_enc_is_term_pass, t1.cpp:1131-1151 */
static int enc_is_term_pass(int numbps, int sty, int bp, int type) {
	if (type == 2 && bp == 0) return 1;
	if (sty & STY_TERMALL) return 1;
	if (sty & STY_LAZY) {
		if (bp == numbps - 4 && type == 2) return 1;
		if (bp < numbps - 4 && type > 0) return 1;
	}
	return 0;
}

/*
 * Tier-1 encode of one block with code-block style switches (t1.cpp:1182-1326):
 *   LAZY 0x01 bypass of the MQ coder for the significance / refinement passes below the fourth plane,
 *   RESET 0x02 context reset after every pass, TERMALL 0x04 every pass terminated, VSC 0x08 stripe-causal
 *   contexts, PTERM 0x10 predictable termination, SEGSYM 0x20 segmentation symbol after each cleanup pass.
 * terms[i] = 1 when pass i ends a codeword segment.  Everything else as gbo_t1_encode_block.
 */
GBO_API int gbo_t1_encode_block_sty(const int32_t *data, int w, int h, int orient, int sty, int do_rd, double wbase,
		uint8_t *out, uint32_t *numbps_out, uint32_t *rates, double *dists, uint8_t *terms, uint64_t *nsym_out) {
	t1e e;
	memset(&e, 0, sizeof(e));
	e.data = data;
	uint32_t mx = 0;
	for (int i = 0; i < w * h; ++i) {
		uint32_t a = (uint32_t) (data[i] < 0 ? -data[i] : data[i]);
		if (a > mx) mx = a;
	}
	int nbits = 0;
	while (mx >> nbits) nbits++;
	int numbps = nbits > 6 ? nbits - 6 : 0;
	*numbps_out = (uint32_t) numbps;
	if (nsym_out) *nsym_out = 0;
	if (numbps == 0)
		return 0;
	if (t1s_alloc(&e.s, w, h, orient))
		return -1;
	e.s.vsc = (sty & STY_VSC) != 0;
	out[-1] = 0;
	mqe_init(&e.q, out);
	int npass = 0;
	double cum = 0.0;
	for (int bp = numbps - 1; bp >= 0; --bp)
		for (int type = (bp == numbps - 1 ? 2 : 0); type < 3; ++type) {
			e.raw = (bp < numbps - 4) && type < 2 && (sty & STY_LAZY);
			if (npass > 0 && terms[npass - 1]) { /* the previous pass closed a segment */
				if (e.raw) mqe_bypass_init(&e.q); else mqe_restart(&e.q);
			}
			if (type == 0) enc_sigpass(&e, bp);
			else if (type == 1) enc_refpass(&e, bp);
			else {
				enc_clnpass(&e, bp);
				if (sty & STY_SEGSYM) { /* mqc_segmark_enc, mqc_enc.cpp:409-413 */
					for (int i = 1; i < 5; ++i) enc_emit(&e, CTX_UNI, i & 1);
				}
			}
			if (do_rd) {
				double x = wbase * (double) (1 << bp);
				x *= x * e.nmsedec / 8192.0;
				cum += x;
				dists[npass] = cum;
			} else
				dists[npass] = 0.0;
			if (enc_is_term_pass(numbps, sty, bp, type)) {
				if (e.raw) mqe_bypass_flush(&e.q, sty & STY_PTERM);
				else if (sty & STY_PTERM) mqe_erterm(&e.q);
				else mqe_flush(&e.q);
				terms[npass] = 1;
				rates[npass] = mqe_numbytes(&e.q);
			} else {
				terms[npass] = 0;
				rates[npass] = mqe_numbytes(&e.q) + (e.raw ? mqe_bypass_extra_bytes(&e.q, sty & STY_PTERM) : (e.q.ct < 5 ? 6u : 5u));
			}
			npass++;
			if (sty & STY_RESET) mq_reset_ctx(e.q.st, e.q.mps);
		}
	uint32_t last = mqe_numbytes(&e.q);
	for (int i = npass - 1; i >= 0; --i) {
		if (rates[i] > last) rates[i] = last; else last = rates[i];
	}
	for (int i = 0; i < npass; ++i)
		if (rates[i] > 0 && out[(int) rates[i] - 1] == 0xFF) rates[i]--;
	if (nsym_out) *nsym_out = e.nsym;
	t1s_free(&e.s);
	return npass;
}

/* ---- decoder ------------------------------------------------------------------------------ */

typedef struct {
	int32_t *data;
	t1s s;
	mqd q;
	int raw;           /* this segment bypasses the arithmetic coder */
	uint32_t rc;       /* raw decoder: current byte, bits left in it (mqc_dec_inl.h:90-112) */
	int rct;
} t1d;

/* mqc_raw_decode, mqc_dec_inl.h:90-112: bit-unstuffing reader; bytes past the segment read as 0xFF */
static int raw_decode(t1d *d) {
	mqd *q = &d->q;
	if (d->rct == 0) {
		uint32_t b = mqd_byte(q, q->pos);
		if (d->rc == 0xFF) {
			if (b > 0x8F) { d->rc = 0xFF; d->rct = 8; }
			else { d->rc = b; q->pos++; d->rct = 7; }
		} else { d->rc = b; q->pos++; d->rct = 8; }
	}
	d->rct--;
	return (int) ((d->rc >> d->rct) & 1);
}

static int dec_get(t1d *d, int cx) { return d->raw ? raw_decode(d) : mqd_decode(&d->q, cx); }

static void dec_sign_and_mark(t1d *d, int x, int y, int32_t oneplushalf) {
	t1s *s = &d->s;
	int xorbit;
	int cx = ctx_sc(s, x, y, &xorbit);
	int neg = dec_get(d, cx);
	if (!d->raw) neg ^= xorbit; /* raw passes carry the sign itself */
	d->data[y * s->w + x] = neg ? -oneplushalf : oneplushalf;
	AT(s->sig, x, y) = 1;
	AT(s->neg, x, y) = (uint8_t) neg;
}

/* t1.cpp:381-441 */
static void dec_sigpass(t1d *d, int bp1) {
	t1s *s = &d->s;
	int32_t one = 1 << bp1, oph = one | (one >> 1);
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x)
			for (int y = y0; y < y0 + 4 && y < s->h; ++y) {
				if (AT(s->sig, x, y) || !any_sig_neighbour(s, x, y))
					continue;
				if (dec_get(d, CTX_ZC0 + ctx_zc(s, x, y)))
					dec_sign_and_mark(d, x, y, oph);
				AT(s->vis, x, y) = 1;
			}
}

/* t1.cpp:476-496, 588-637 */
static void dec_refpass(t1d *d, int bp1) {
	t1s *s = &d->s;
	int32_t poshalf = (1 << bp1) >> 1;
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x)
			for (int y = y0; y < y0 + 4 && y < s->h; ++y) {
				if (!AT(s->sig, x, y) || AT(s->vis, x, y))
					continue;
				int bit = dec_get(d, ctx_mr(s, x, y));
				int32_t *p = &d->data[y * s->w + x];
				*p += (bit ^ (*p < 0)) ? poshalf : -poshalf;
				AT(s->refd, x, y) = 1;
			}
}

/* t1.cpp:784-870 */
static void dec_clnpass(t1d *d, int bp1) {
	t1s *s = &d->s;
	int32_t one = 1 << bp1, oph = one | (one >> 1);
	for (int y0 = 0; y0 < s->h; y0 += 4)
		for (int x = 0; x < s->w; ++x) {
			int y = y0;
			int runmode = (y0 + 4 <= s->h);
			for (int k = 0; k < 4 && runmode; ++k)
				if (AT(s->sig, x, y0 + k) || AT(s->vis, x, y0 + k) || any_sig_neighbour(s, x, y0 + k))
					runmode = 0;
			if (runmode) {
				if (!mqd_decode(&d->q, CTX_AGG))
					continue;
				int r = mqd_decode(&d->q, CTX_UNI);
				r = (r << 1) | mqd_decode(&d->q, CTX_UNI);
				y = y0 + r;
				dec_sign_and_mark(d, x, y, oph);
				y++;
			}
			for (; y < y0 + 4 && y < s->h; ++y) {
				if (AT(s->sig, x, y) || AT(s->vis, x, y))
					continue;
				if (mqd_decode(&d->q, CTX_ZC0 + ctx_zc(s, x, y)))
					dec_sign_and_mark(d, x, y, oph);
			}
		}
	memset(s->vis, 0, (size_t) s->S * (s->h + 2));
}

/*
 * Tier-1 decode of one single-segment block (t1.cpp:1038-1130 with cblk_sty == 0, roishift == 0).
 * out = w*h values carrying one extra low bit (2x the mid-point reconstruction), zero where
 * nothing was decoded.
 */
GBO_API int gbo_t1_decode_block(const uint8_t *bytes, uint32_t len, int numpasses, int numbps, int orient,
		int w, int h, int32_t *out) {
	t1d d;
	memset(&d, 0, sizeof(d));
	memset(out, 0, sizeof(int32_t) * (size_t) w * h);
	if (numbps >= 31)
		return 1;
	if (t1s_alloc(&d.s, w, h, orient))
		return -1;
	d.data = out;
	mqd_init(&d.q, bytes, len);
	int bp1 = numbps, type = 2;
	for (int p = 0; p < numpasses && bp1 >= 1; ++p) {
		if (type == 0) dec_sigpass(&d, bp1);
		else if (type == 1) dec_refpass(&d, bp1);
		else dec_clnpass(&d, bp1);
		if (++type == 3) { type = 0; bp1--; }
	}
	t1s_free(&d.s);
	return 0;
}

/*
 * Tier-1 decode of one block from its codeword segments, with code-block style switches (t1.cpp:1038-1130).
 * bytes = the segments back to back; seg_len[i] / seg_passes[i] as Tier-2 delivers them (T2.cpp:835-851: one pass per
 * segment with TERMALL; 10, then 2, 1, 2, 1 ... with LAZY; otherwise a single segment).
 */
GBO_API int gbo_t1_decode_block_roi(const uint8_t *bytes, const uint32_t *seg_len, const uint32_t *seg_passes, int nsegs,
		int numbps, int roishift, int orient, int sty, int w, int h, int32_t *out);

GBO_API int gbo_t1_decode_block_segs(const uint8_t *bytes, const uint32_t *seg_len, const uint32_t *seg_passes, int nsegs,
		int numbps, int orient, int sty, int w, int h, int32_t *out) {
	return gbo_t1_decode_block_roi(bytes, seg_len, seg_passes, nsegs, numbps, 0, orient, sty, w, h, out);
}

/* the same for a component with a max-shift region of interest: numbps = cblk->numbps - roishift as T1Part1::decode
 * passes it (T1Part1.cpp:184-186); decoding starts at plane roishift + numbps (t1.cpp:1055) and samples at or above
 * 2^roishift are shifted back down afterwards (T1Part1::post_decode, T1Part1.cpp:230-252) */
GBO_API int gbo_t1_decode_block_roi(const uint8_t *bytes, const uint32_t *seg_len, const uint32_t *seg_passes, int nsegs,
		int numbps, int roishift, int orient, int sty, int w, int h, int32_t *out) {
	t1d d;
	memset(&d, 0, sizeof(d));
	memset(out, 0, sizeof(int32_t) * (size_t) w * h);
	if (numbps + roishift >= 31)
		return 1;
	if (t1s_alloc(&d.s, w, h, orient))
		return -1;
	d.s.vsc = (sty & STY_VSC) != 0;
	d.data = out;
	mq_reset_ctx(d.q.st, d.q.mps);
	int bp1 = numbps + roishift, type = 2;
	uint32_t off = 0;
	for (int sg = 0; sg < nsegs; ++sg) {
		d.raw = (bp1 <= numbps - 4) && type < 2 && (sty & STY_LAZY);
		uint8_t st[NCTX], mps[NCTX];
		memcpy(st, d.q.st, NCTX);
		memcpy(mps, d.q.mps, NCTX);
		mqd_init(&d.q, bytes + off, seg_len[sg]); /* resets the contexts: put them back, only RESET clears them */
		memcpy(d.q.st, st, NCTX);
		memcpy(d.q.mps, mps, NCTX);
		if (d.raw) { d.q.pos = 0; d.rc = 0; d.rct = 0; } /* mqc_raw_init_dec, mqc_dec.cpp:195-200 */
		off += seg_len[sg];
		for (uint32_t p = 0; p < seg_passes[sg] && bp1 >= 1; ++p) {
			if (type == 0) dec_sigpass(&d, bp1);
			else if (type == 1) dec_refpass(&d, bp1);
			else {
				dec_clnpass(&d, bp1);
				if (sty & STY_SEGSYM) /* t1_dec_clnpass_check_segsym, t1.cpp:870-888: read, a mismatch only warns */
					for (int i = 0; i < 4; ++i) mqd_decode(&d.q, CTX_UNI);
			}
			if ((sty & STY_RESET) && !d.raw) mq_reset_ctx(d.q.st, d.q.mps);
			if (++type == 3) { type = 0; bp1--; }
		}
	}
	if (roishift) {
		const int32_t thresh = 1 << roishift;
		for (int i = 0; i < w * h; ++i) {
			int32_t mag = out[i] < 0 ? -out[i] : out[i];
			if (mag >= thresh) out[i] = out[i] < 0 ? -(mag >> roishift) : (mag >> roishift);
		}
	}
	t1s_free(&d.s);
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Tile-component geometry (TileComponent.cpp:193-489, Tier1.cpp:53-86)                        */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
	uint32_t resno, orient, precno, cblkno; /* traversal position */
	uint32_t x0, y0, x1, y1;                /* band coordinates */
	uint32_t off_x, off_y;                  /* position inside the Mallat-layout tile plane */
} gbo_block;

static inline uint32_t fdiv2n(uint32_t a, uint32_t n) { return a >> n; }
static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* Enumerates the code blocks of one tile-component in the host's order (resno, band, precinct,
 * block).  prc_expn[2*resno+{0,1}] = precinct width/height exponents.  Returns the count; if
 * `out` is NULL only counts. */
GBO_API int gbo_enumerate_blocks(uint32_t tx0, uint32_t ty0, uint32_t tx1, uint32_t ty1, uint32_t numres,
		uint32_t cblkw_expn, uint32_t cblkh_expn, const uint32_t *prc_expn, gbo_block *out) {
	int n = 0;
	for (uint32_t resno = 0; resno < numres; ++resno) {
		uint32_t lvl = numres - 1 - resno;
		uint32_t rx0 = cdiv2n(tx0, lvl), ry0 = cdiv2n(ty0, lvl), rx1 = cdiv2n(tx1, lvl), ry1 = cdiv2n(ty1, lvl);
		uint32_t pdx = prc_expn[2 * resno], pdy = prc_expn[2 * resno + 1];
		uint32_t px0 = fdiv2n(rx0, pdx) << pdx, py0 = fdiv2n(ry0, pdy) << pdy;
		uint32_t px1 = cdiv2n(rx1, pdx) << pdx, py1 = cdiv2n(ry1, pdy) << pdy;
		uint32_t pw = rx0 == rx1 ? 0 : (px1 - px0) >> pdx, ph = ry0 == ry1 ? 0 : (py1 - py0) >> pdy;
		uint32_t gx0, gy0, gwe, ghe, nbands;
		if (resno == 0) { gx0 = px0; gy0 = py0; gwe = pdx; ghe = pdy; nbands = 1; }
		else { gx0 = cdiv2n(px0, 1); gy0 = cdiv2n(py0, 1); gwe = pdx - 1; ghe = pdy - 1; nbands = 3; }
		uint32_t cwe = umin(cblkw_expn, gwe), che = umin(cblkh_expn, ghe);
		/* size of the next-lower resolution: offset of the HL/LH/HH quadrants */
		uint32_t lw = resno ? cdiv2n(tx1, lvl + 1) - cdiv2n(tx0, lvl + 1) : 0;
		uint32_t lh = resno ? cdiv2n(ty1, lvl + 1) - cdiv2n(ty0, lvl + 1) : 0;
		for (uint32_t b = 0; b < nbands; ++b) {
			uint32_t orient = resno == 0 ? 0 : b + 1;
			uint32_t bx0, by0, bx1, by1;
			if (resno == 0) { bx0 = rx0; by0 = ry0; bx1 = rx1; by1 = ry1; }
			else {
				uint64_t xo = (uint64_t) (orient & 1) << lvl, yo = (uint64_t) (orient >> 1) << lvl;
				bx0 = (uint32_t) (((uint64_t) tx0 - xo + ((1ull << (lvl + 1)) - 1)) >> (lvl + 1));
				by0 = (uint32_t) (((uint64_t) ty0 - yo + ((1ull << (lvl + 1)) - 1)) >> (lvl + 1));
				bx1 = (uint32_t) (((uint64_t) tx1 - xo + ((1ull << (lvl + 1)) - 1)) >> (lvl + 1));
				by1 = (uint32_t) (((uint64_t) ty1 - yo + ((1ull << (lvl + 1)) - 1)) >> (lvl + 1));
			}
			for (uint32_t p = 0; p < pw * ph; ++p) {
				uint32_t cx0 = gx0 + (p % pw) * (1u << gwe), cy0 = gy0 + (p / pw) * (1u << ghe);
				uint32_t qx0 = umax(cx0, bx0), qy0 = umax(cy0, by0);
				uint32_t qx1 = umin(cx0 + (1u << gwe), bx1), qy1 = umin(cy0 + (1u << ghe), by1);
				if (qx1 < qx0 || qy1 < qy0) /* zero-area blocks of an empty band stay (TileComponent.cpp:384-404) */
					continue;
				uint32_t kx0 = fdiv2n(qx0, cwe) << cwe, ky0 = fdiv2n(qy0, che) << che;
				uint32_t kx1 = cdiv2n(qx1, cwe) << cwe, ky1 = cdiv2n(qy1, che) << che;
				uint32_t cw = (kx1 - kx0) >> cwe, ch = (ky1 - ky0) >> che;
				for (uint32_t k = 0; k < cw * ch; ++k) {
					uint32_t ax = kx0 + (k % cw) * (1u << cwe), ay = ky0 + (k / cw) * (1u << che);
					if (out) {
						gbo_block *o = &out[n];
						o->resno = resno; o->orient = orient; o->precno = p; o->cblkno = k;
						o->x0 = umax(ax, qx0); o->y0 = umax(ay, qy0);
						o->x1 = umin(ax + (1u << cwe), qx1); o->y1 = umin(ay + (1u << che), qy1);
						o->off_x = o->x0 - bx0 + ((orient & 1) ? lw : 0);
						o->off_y = o->y0 - by0 + ((orient & 2) ? lh : 0);
					}
					n++;
				}
			}
		}
	}
	return n;
}

/* ---- PCRD preparation: feasible truncation points of one code block -----------------------------------
 * RateControl.cpp:31-118 (convexHull) and :159-168 (slopeToLog).  Walks the passes of a block; a pass is a
 * feasible truncation point if the distortion-rate slope towards every earlier feasible point is positive,
 * finite and strictly decreasing; its slope is stored as ln(slope) in 8.8 fixed point (0 = not feasible).
 * len[p]: bytes the pass adds; dist[p]: CUMULATIVE distortion decrease up to and including pass p. */
static uint16_t slope_to_log(double slope) {
	const double cutoff = pow(2, 64), scale = 256 / log(2), shift = 1 << 16;
	double v;
	if (slope > cutoff) slope = cutoff;
	v = log(slope) * scale - log(cutoff) * scale + shift;
	if (v < 1) v = 1;
	if (v > 0xFFFF) v = 0xFFFF;
	return (uint16_t) v;
}

GBO_API void gbo_rd_convex_hull(const uint32_t *len, const double *dist, uint32_t numpasses, uint16_t *slope) {
	double *cache = (double*) malloc(sizeof(double) * (numpasses ? numpasses : 1));
	uint32_t p;
	for (p = 0; p < numpasses; ++p) {
		double dd = 0, dr = 0;
		int q = (int) p; /* the intermediate point, walking down from p */
		slope[p] = 0;
		for (;;) {
			dr += len[q];
			dd += q == 0 ? dist[q] : dist[q] - dist[q - 1];
			if (dd <= 0) { slope[p] = 0; break; } /* Corollary 8.3: every intermediate slope must be positive */
			--q;
			if (q == -1) { cache[p] = dd / dr; slope[p] = slope_to_log(cache[p]); break; }
			if (slope[q] == 0) continue; /* rejected earlier */
			if (dr == 0) slope[q] = 0; /* the slope must be finite */
			else if (cache[q] * dr <= dd) slope[q] = 0; /* ... and strictly decreasing */
			else {
				cache[p] = dd / dr;
				slope[p] = slope_to_log(cache[p]);
				if (slope[p] >= slope[q]) slope[q] = 0; /* the coarser log domain may break monotonicity: drop the earlier point */
				break;
			}
		}
	}
	free(cache);
}
