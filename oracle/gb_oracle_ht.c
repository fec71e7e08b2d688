/*
 * gb_oracle_ht.c -- TEST INFRASTRUCTURE (part of oracle/libgb_oracle.so).  CPU restatement of the HTJ2K block coder the
 * reference uses when a component's code-block style has the HT bit (grk_compress -M 64): the CLEANUP pass of Rec. ITU-T
 * T.814 (ISO/IEC 15444-15), which is all the reference's encoder emits (T1HT.cpp:104-133: one pass, numbps = 1).
 *
 *   T1HT::preEncode        t1/t1_ht/T1HT.cpp:56-103        sign-magnitude, MSB aligned          gbo_ht_quantise_block
 *   ojph_encode_codeblock  coding/ojph_block_encoder.cpp:465-938  MagSgn + MEL + VLC byte streams  gbo_ht_encode_block
 *   ojph_decode_codeblock  coding/ojph_block_decoder.cpp:687-1200 cleanup pass only                gbo_ht_decode_block
 *   T1HT::postDecode       t1/t1_ht/T1HT.cpp:176-251       back to the tile buffer              gbo_ht_dequantise_block
 *
 * Written from the structure of the standard, one sample / one quad at a time, with byte-wise bit readers and writers
 * (the reference moves 32 bits at a time); pinned bit-for-bit against the live reference by tests/test_oracle_ht.py.
 * Only tests/, __graft_entry__.smoke() and the CPU-baseline legs of bench.py may call into this library.
 *
 * Terms (T.814): a quad is a 2x2 group of samples scanned column by column (0 top-left, 1 bottom-left, 2 top-right,
 * 3 bottom-right); rho = significance pattern of the quad; U = bit count of the quad's largest magnitude exponent;
 * kappa = the exponent predicted from the row of quads above; u = U - kappa is sent with the U-VLC code; eps marks the
 * samples whose exponent reaches U (the "exponent max bound" pattern that selects the CxtVLC codeword).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__GNUC__)
#define GBO_API __attribute__((visibility("default")))
#else
#define GBO_API
#endif

#include "ht_tables.inc"

static inline int bits_of(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* ---- quantisation into the block coder's sign-magnitude form (T1HT.cpp:56-103) ---------------------------------- */
GBO_API uint32_t gbo_ht_quantise_block(const int32_t *src, uint32_t stride, uint32_t w, uint32_t h, int reversible, float stepsize,
		uint32_t k_msbs, int32_t *out) {
	uint32_t maximum = 0;
	for (uint32_t y = 0; y < h; ++y)
		for (uint32_t x = 0; x < w; ++x) {
			const int32_t t = src[(size_t) y * stride + x];
			int32_t res;
			if (reversible) {
				const int32_t shift = 31 - ((int32_t) k_msbs + 1);
				const int32_t val = t >= 0 ? t : -t;
				res = (int32_t) ((t >= 0 ? 0u : 0x80000000u) | ((uint32_t) val << shift));
				if ((uint32_t) res > maximum) maximum = (uint32_t) res;
			} else {
				const int32_t shift = 31 - ((int32_t) k_msbs + 1) - 11;
				const float inv = 1.0f / stepsize; /* Tier1.cpp:78 */
				const int32_t q = (int32_t) ((float) t * inv * (float) (1 << shift)); /* left to right, truncation */
				const int32_t val = q >= 0 ? q : -q;
				if ((uint32_t) val > maximum) maximum = (uint32_t) val;
				res = (int32_t) ((q >= 0 ? 0u : 0x80000000u) | (uint32_t) val);
			}
			out[(size_t) y * w + x] = res;
		}
	return maximum;
}

/* T1HT.cpp:211-236: reversible: magnitude >> (31 - (k_msbs + 1)); irreversible: (float) magnitude * stepsize */
GBO_API void gbo_ht_dequantise_block(const int32_t *dec, uint32_t w, uint32_t h, int reversible, float stepsize, uint32_t k_msbs,
		int32_t *dst, uint32_t dst_stride) {
	for (uint32_t y = 0; y < h; ++y)
		for (uint32_t x = 0; x < w; ++x) {
			const int32_t t = dec[(size_t) y * w + x];
			const int32_t mag = t & 0x7FFFFFFF;
			if (reversible) {
				const int32_t val = mag >> (31 - ((int32_t) k_msbs + 1));
				dst[(size_t) y * dst_stride + x] = (t & (int32_t) 0x80000000) ? -val : val;
			} else {
				const float val = (float) mag * stepsize;
				const float r = (t & (int32_t) 0x80000000) ? -val : val;
				memcpy(&dst[(size_t) y * dst_stride + x], &r, 4);
			}
		}
}

/* ================================================================================================================ */
/* encoder                                                                                                          */
/* ================================================================================================================ */

/* forward writers: MagSgn packs bits LSB first, MEL packs MSB first; the byte after a 0xFF carries 7 bits */
typedef struct { uint8_t *buf; int pos, cap, used, limit; uint32_t acc; } fwd_writer;

static void ms_put(fwd_writer *s, uint32_t bits, int n) {
	while (n > 0) {
		int take = s->limit - s->used;
		if (take > n) take = n;
		s->acc |= (bits & ((1u << take) - 1u)) << s->used;
		s->used += take;
		bits >>= take;
		n -= take;
		if (s->used == s->limit) {
			if (s->pos < s->cap) s->buf[s->pos] = (uint8_t) s->acc;
			s->pos++;
			s->limit = s->acc == 0xFFu ? 7 : 8;
			s->acc = 0;
			s->used = 0;
		}
	}
}

static void ms_finish(fwd_writer *s) {
	if (s->used) { /* pad the open byte with ones; a padded 0xFF is dropped */
		const int pad = s->limit - s->used;
		s->acc |= ((1u << pad) - 1u) << s->used;
		if (s->acc != 0xFFu) {
			if (s->pos < s->cap) s->buf[s->pos] = (uint8_t) s->acc;
			s->pos++;
		}
	} else if (s->limit == 7) s->pos--; /* the stream may not end with 0xFF */
}

typedef struct { fwd_writer w; int left; int run, k, threshold; } mel_writer;
static const int MEL_E[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5};

static void mel_bit(mel_writer *m, int v) {
	m->w.acc = (m->w.acc << 1) | (uint32_t) v;
	if (--m->left == 0) {
		if (m->w.pos < m->w.cap) m->w.buf[m->w.pos] = (uint8_t) m->w.acc;
		m->w.pos++;
		m->left = m->w.acc == 0xFFu ? 7 : 8;
		m->w.acc = 0;
	}
}

/* one MEL event: 0 extends the run (a 1 bit is sent when the run reaches the threshold of the state), 1 ends it */
static void mel_event(mel_writer *m, int one) {
	if (!one) {
		if (++m->run >= m->threshold) {
			mel_bit(m, 1);
			m->run = 0;
			if (m->k < 12) m->k++;
			m->threshold = 1 << MEL_E[m->k];
		}
	} else {
		mel_bit(m, 0);
		for (int t = MEL_E[m->k]; t > 0;) mel_bit(m, (m->run >> --t) & 1);
		m->run = 0;
		if (m->k > 0) m->k--;
		m->threshold = 1 << MEL_E[m->k];
	}
}

/* VLC stream: grows downwards from the end of its buffer, bits LSB first; a byte that follows one > 0x8F and whose low
 * seven bits are all ones keeps its MSB clear */
typedef struct { uint8_t *end; int pos, cap, used; uint32_t acc; int prev_gt_8f; } rev_writer;

static void vlc_put(rev_writer *s, uint32_t bits, int n) {
	while (n > 0) {
		int room = 8 - s->prev_gt_8f - s->used;
		int take = room < n ? room : n;
		s->acc |= (bits & ((1u << take) - 1u)) << s->used;
		s->used += take;
		room -= take;
		n -= take;
		bits >>= take;
		if (room == 0) {
			if (s->prev_gt_8f && s->acc != 0x7Fu) { s->prev_gt_8f = 0; continue; } /* the eighth bit is usable after all */
			if (s->pos < s->cap) s->end[-s->pos] = (uint8_t) s->acc;
			s->pos++;
			s->prev_gt_8f = s->acc > 0x8Fu;
			s->acc = 0;
			s->used = 0;
		}
	}
}

/* U-VLC code of u (T.814 Table 3): prefix 1 / 01 / 001 / 000, then 0, 0, 1 or 5 suffix bits */
static void uvlc_parts(int u, uint32_t *pre, int *pre_len, uint32_t *suf, int *suf_len) {
	if (u == 0) { *pre = 0; *pre_len = 0; *suf = 0; *suf_len = 0; }
	else if (u == 1) { *pre = 1; *pre_len = 1; *suf = 0; *suf_len = 0; }
	else if (u == 2) { *pre = 2; *pre_len = 2; *suf = 0; *suf_len = 0; }
	else if (u <= 4) { *pre = 4; *pre_len = 3; *suf = (uint32_t) (u - 3); *suf_len = 1; }
	else { *pre = 0; *pre_len = 3; *suf = (uint32_t) (u - 5); *suf_len = 5; }
}

typedef struct { int rho, emax, e[4]; uint32_t v[4]; } ht_quad;

/* the four samples of the quad whose top-left corner is (x, y) */
static void load_quad(const int32_t *buf, int stride, int w, int h, int x, int y, int p, ht_quad *q) {
	memset(q, 0, sizeof(*q));
	for (int i = 0; i < 4; ++i) {
		const int xx = x + (i >> 1), yy = y + (i & 1);
		if (xx >= w || yy >= h) continue;
		const uint32_t t = (uint32_t) buf[(size_t) yy * stride + xx];
		uint32_t val = (t + t) >> p; /* drops the sign: 2 * mu_p + a lower bit */
		val &= ~1u;
		if (!val) continue;
		q->rho |= 1 << i;
		q->e[i] = bits_of(val - 1);       /* bit count of 2 mu_p - 1 */
		if (q->e[i] > q->emax) q->emax = q->e[i];
		q->v[i] = (val - 2) + (t >> 31);  /* 2 (mu_p - 1) + sign */
	}
}

/* Cleanup pass of one code block.  sm: sign-magnitude samples (bit 31 = sign, magnitude MSB aligned so that the top
 * coded bit plane sits at bit 30 - missing_msbs).  Returns the number of bytes written to out, or -1 if cap is too small. */
GBO_API int gbo_ht_encode_block(const int32_t *sm, int w, int h, int stride, int missing_msbs, uint8_t *out, int cap) {
	const int p = 30 - missing_msbs;
	enum { MS_CAP = 65536, MEL_CAP = 1024, VLC_CAP = 16384 };
	uint8_t *ms_buf = (uint8_t*) malloc(MS_CAP), *mel_buf = (uint8_t*) malloc(MEL_CAP), *vlc_buf = (uint8_t*) malloc(VLC_CAP);
	fwd_writer ms = {ms_buf, 0, MS_CAP, 0, 8, 0};
	mel_writer mel = {{mel_buf, 0, MEL_CAP, 0, 8, 0}, 8, 0, 0, 1};
	rev_writer vlc = {vlc_buf + VLC_CAP - 1, 1, VLC_CAP, 4, 0xF, 1};
	vlc.end[0] = 0xFF;
	const int nq = (w + 1) / 2; /* quads per row */
	/* line state for the row of quads below: exponent and significance of every bottom-row sample of the quads above */
	uint8_t *e_bot = (uint8_t*) calloc((size_t) 2 * nq + 4, 1), *s_bot = (uint8_t*) calloc((size_t) 2 * nq + 4, 1);
	uint8_t *e_new = (uint8_t*) calloc((size_t) 2 * nq + 4, 1), *s_new = (uint8_t*) calloc((size_t) 2 * nq + 4, 1);
	for (int y = 0; y < h; y += 2) {
		const int first = y == 0;
		const uint16_t *tbl = first ? HT_VLC_ENC0 : HT_VLC_ENC1;
		memset(e_new, 0, (size_t) 2 * nq + 4);
		memset(s_new, 0, (size_t) 2 * nq + 4);
		int prev_rho = 0;
		for (int qx = 0; qx < nq; qx += 2) { /* quads are coded in pairs */
			ht_quad Q[2];
			int U[2] = {0, 0}, u[2] = {0, 0}, present[2] = {1, qx + 1 < nq};
			uint16_t tuple[2] = {0, 0};
			for (int k = 0; k < 2; ++k) {
				if (!present[k]) break;
				const int q = qx + k;
				load_quad(sm, stride, w, h, 2 * q, y, p, &Q[k]);
				int cq, kappa = 1;
				if (first) cq = (prev_rho >> 1) | (prev_rho & 1);
				else {
					/* column index of the quad's left sample in the row above: 2q; neighbours nw = 2q-1, n = 2q, ne = 2q+1, nf = 2q+2 */
					const int snw = q ? s_bot[2 * q - 1] : 0, sn = s_bot[2 * q], sne = s_bot[2 * q + 1], snf = s_bot[2 * q + 2];
					cq = (snw | sn) | (((prev_rho >> 2) | (prev_rho >> 3)) & 1) << 1 | (sne | snf) << 2;
					if (Q[k].rho & (Q[k].rho - 1)) { /* more than one significant sample: predict from the exponents above */
						int emax = q ? e_bot[2 * q - 1] : 0;
						if (e_bot[2 * q] > emax) emax = e_bot[2 * q];
						if (e_bot[2 * q + 1] > emax) emax = e_bot[2 * q + 1];
						if (e_bot[2 * q + 2] > emax) emax = e_bot[2 * q + 2];
						kappa = emax - 1 > 1 ? emax - 1 : 1;
					}
				}
				U[k] = Q[k].emax > kappa ? Q[k].emax : kappa;
				u[k] = U[k] - kappa;
				int eps = 0;
				if (u[k] > 0)
					for (int i = 0; i < 4; ++i) eps |= (Q[k].e[i] == Q[k].emax) << i;
				tuple[k] = tbl[(cq << 8) | (Q[k].rho << 4) | eps];
				vlc_put(&vlc, tuple[k] >> 8, (tuple[k] >> 4) & 7);
				if (cq == 0) mel_event(&mel, Q[k].rho != 0);
				for (int i = 0; i < 4; ++i) {
					const int m = (Q[k].rho >> i & 1) ? U[k] - (tuple[k] >> i & 1) : 0;
					ms_put(&ms, Q[k].v[i] & ((1u << m) - 1u), m);
				}
				e_new[2 * q] = (uint8_t) Q[k].e[1]; e_new[2 * q + 1] = (uint8_t) Q[k].e[3];
				s_new[2 * q] = (uint8_t) (Q[k].rho >> 1 & 1); s_new[2 * q + 1] = (uint8_t) (Q[k].rho >> 3 & 1);
				prev_rho = Q[k].rho;
			}
			/* the U-VLC codes of the pair: both prefixes, then both suffixes */
			uint32_t pre[2], suf[2];
			int pl[2], sl[2];
			if (first) {
				if (u[0] > 0 && u[1] > 0) mel_event(&mel, (u[0] < u[1] ? u[0] : u[1]) > 2);
				if (u[0] > 2 && u[1] > 2) {
					uvlc_parts(u[0] - 2, &pre[0], &pl[0], &suf[0], &sl[0]);
					uvlc_parts(u[1] - 2, &pre[1], &pl[1], &suf[1], &sl[1]);
					vlc_put(&vlc, pre[0], pl[0]); vlc_put(&vlc, pre[1], pl[1]);
					vlc_put(&vlc, suf[0], sl[0]); vlc_put(&vlc, suf[1], sl[1]);
				} else if (u[0] > 2 && u[1] > 0) {
					uvlc_parts(u[0], &pre[0], &pl[0], &suf[0], &sl[0]);
					vlc_put(&vlc, pre[0], pl[0]);
					vlc_put(&vlc, (uint32_t) (u[1] - 1), 1);
					vlc_put(&vlc, suf[0], sl[0]);
				} else {
					uvlc_parts(u[0], &pre[0], &pl[0], &suf[0], &sl[0]);
					uvlc_parts(u[1], &pre[1], &pl[1], &suf[1], &sl[1]);
					vlc_put(&vlc, pre[0], pl[0]); vlc_put(&vlc, pre[1], pl[1]);
					vlc_put(&vlc, suf[0], sl[0]); vlc_put(&vlc, suf[1], sl[1]);
				}
			} else {
				uvlc_parts(u[0], &pre[0], &pl[0], &suf[0], &sl[0]);
				uvlc_parts(u[1], &pre[1], &pl[1], &suf[1], &sl[1]);
				vlc_put(&vlc, pre[0], pl[0]); vlc_put(&vlc, pre[1], pl[1]);
				vlc_put(&vlc, suf[0], sl[0]); vlc_put(&vlc, suf[1], sl[1]);
			}
		}
		uint8_t *t = e_bot; e_bot = e_new; e_new = t;
		t = s_bot; s_bot = s_new; s_new = t;
	}
	/* ---- termination: the open MEL and VLC bytes are fused into one when their used bits do not collide ---- */
	if (mel.run > 0) mel_bit(&mel, 1);
	{
		const uint32_t mel_tmp = (mel.w.acc << mel.left) & 0xFFu;
		const uint32_t mel_mask = (0xFFu << mel.left) & 0xFFu, vlc_mask = 0xFFu >> (8 - vlc.used);
		if ((mel_mask | vlc_mask) != 0) {
			const uint32_t fuse = mel_tmp | vlc.acc;
			if ((((fuse ^ mel_tmp) & mel_mask) | ((fuse ^ vlc.acc) & vlc_mask)) == 0 && fuse != 0xFFu && vlc.pos > 1) {
				if (mel.w.pos < MEL_CAP) mel_buf[mel.w.pos] = (uint8_t) fuse;
				mel.w.pos++;
			} else {
				if (mel.w.pos < MEL_CAP) mel_buf[mel.w.pos] = (uint8_t) mel_tmp;
				mel.w.pos++;
				if (vlc.pos < VLC_CAP) vlc.end[-vlc.pos] = (uint8_t) vlc.acc;
				vlc.pos++;
			}
		}
	}
	ms_finish(&ms);
	const int total = ms.pos + mel.w.pos + vlc.pos;
	int rc = -1;
	if (total <= cap && ms.pos <= MS_CAP && mel.w.pos <= MEL_CAP && vlc.pos <= VLC_CAP) {
		memcpy(out, ms_buf, (size_t) ms.pos);
		memcpy(out + ms.pos, mel_buf, (size_t) mel.w.pos);
		memcpy(out + ms.pos + mel.w.pos, vlc.end - vlc.pos + 1, (size_t) vlc.pos);
		/* the last twelve bits locate the MEL + VLC suffix (Scup) */
		const int scup = mel.w.pos + vlc.pos;
		out[total - 1] = (uint8_t) (scup >> 4);
		out[total - 2] = (uint8_t) ((out[total - 2] & 0xF0) | (scup & 0xF));
		rc = total;
	}
	free(ms_buf); free(mel_buf); free(vlc_buf); free(e_bot); free(s_bot); free(e_new); free(s_new);
	return rc;
}

/* ================================================================================================================ */
/* decoder (cleanup pass)                                                                                           */
/* ================================================================================================================ */

/* MagSgn: forward, LSB first, bytes past the end read as 0xFF; after a 0xFF the next byte gives 7 bits */
typedef struct { const uint8_t *p; int size, pos; uint64_t acc; int bits, unstuff; } fwd_reader;

static void fwd_fill(fwd_reader *r) {
	while (r->bits <= 56) {
		const uint32_t d = r->pos < r->size ? r->p[r->pos] : 0xFFu;
		r->pos++;
		r->acc |= (uint64_t) d << r->bits;
		r->bits += 8 - r->unstuff;
		r->unstuff = d == 0xFFu;
	}
}
static uint32_t fwd_peek(fwd_reader *r) { if (r->bits < 32) fwd_fill(r); return (uint32_t) r->acc; }
static void fwd_skip(fwd_reader *r, int n) { r->acc >>= n; r->bits -= n; }

/* MEL: forward, MSB first; the last byte of the segment has its low nibble forced to ones, bytes past it read as 0xFF */
typedef struct { const uint8_t *p; int size, pos; uint64_t acc; int bits, unstuff; int k, run, one; } mel_reader;

static int mel_next_bit(mel_reader *m) {
	if (m->bits == 0) {
		uint32_t d = m->pos < m->size ? m->p[m->pos] : 0xFFu;
		if (m->pos == m->size - 1) d |= 0xFu;
		m->pos++;
		const int n = 8 - m->unstuff;
		m->acc = d & ((1u << n) - 1u); /* a stuffed byte contributes its low seven bits */
		m->bits = n;
		m->unstuff = d == 0xFFu;
	}
	m->bits--;
	return (int) (m->acc >> m->bits) & 1;
}

/* next MEL event: 1 = the quad (or the U-VLC pair condition) is "on" */
static int mel_event_read(mel_reader *m) {
	if (m->run == 0 && !m->one) {
		const int e = MEL_E[m->k];
		if (mel_next_bit(m)) { m->run = 1 << e; m->one = 0; if (m->k < 12) m->k++; }
		else {
			int r = 0;
			for (int i = 0; i < e; ++i) r = (r << 1) | mel_next_bit(m);
			m->run = r; m->one = 1;
			if (m->k > 0) m->k--;
		}
	}
	if (m->run > 0) { m->run--; return 0; }
	m->one = 0;
	return 1;
}

/* VLC: backwards from the byte before the last one, LSB first; a byte <= 0x8F... see vlc_put for the stuffing rule */
typedef struct { const uint8_t *base; int pos; uint64_t acc; int bits, unstuff; } rev_reader;

static void rev_fill(rev_reader *r) {
	while (r->bits <= 56) {
		const uint32_t d = r->pos >= 0 ? r->base[r->pos] : 0u;
		r->pos--;
		const int n = 8 - ((r->unstuff && (d & 0x7Fu) == 0x7Fu) ? 1 : 0);
		r->acc |= (uint64_t) d << r->bits; /* a stuffed byte has its MSB clear, so nothing leaks into the next one */
		r->bits += n;
		r->unstuff = d > 0x8Fu;
	}
}
static uint32_t rev_peek(rev_reader *r) { if (r->bits < 32) rev_fill(r); return (uint32_t) r->acc; }
static void rev_skip(rev_reader *r, int n) { r->acc >>= n; r->bits -= n; }

/* one U-VLC prefix: returns the prefix value (1, 2, 3 or 5 = long) and consumes its bits */
static int uvlc_prefix(uint32_t *v, int *used) {
	int pv, pl;
	if (*v & 1) { pv = 1; pl = 1; } else if (*v & 2) { pv = 2; pl = 2; } else if (*v & 4) { pv = 3; pl = 3; } else { pv = 5; pl = 3; }
	*v >>= pl; *used += pl;
	return pv;
}
static int uvlc_suffix(int prefix, uint32_t *v, int *used) {
	const int sl = prefix == 3 ? 1 : prefix == 5 ? 5 : 0;
	const int s = (int) (*v & ((1u << sl) - 1u));
	*v >>= sl; *used += sl;
	return prefix + s;
}

/* Decodes the cleanup pass into sign-magnitude samples (bit 31 = sign; magnitude with the reconstruction half bit set, top
 * coded plane at bit 30 - missing_msbs).  Returns 0, or 1 when the suffix length is inconsistent (nothing is written). */
GBO_API int gbo_ht_decode_block(const uint8_t *data, int lcup, int missing_msbs, int w, int h, int stride, int32_t *out) {
	const int p = 30 - missing_msbs;
	if (lcup < 2) return 1;
	const int scup = ((int) data[lcup - 1] << 4) + (data[lcup - 2] & 0xF);
	if (scup > lcup || scup < 2) return 1;
	fwd_reader ms = {data, lcup - scup, 0, 0, 0, 0};
	mel_reader mel = {data + lcup - scup, scup - 1, 0, 0, 0, 0, 0, 0, 0};
	rev_reader vlc = {data, lcup - 3, 0, 0, 0};
	{ /* the upper nibble of the byte that also holds the low nibble of Scup opens the VLC stream */
		const uint32_t d = data[lcup - 2];
		vlc.acc = d >> 4;
		vlc.bits = 4 - ((vlc.acc & 7) == 7);
		vlc.unstuff = (d | 0xF) > 0x8F;
	}
	const int nq = (w + 1) / 2;
	uint8_t *e_bot = (uint8_t*) calloc((size_t) 2 * nq + 4, 1), *s_bot = (uint8_t*) calloc((size_t) 2 * nq + 4, 1);
	uint8_t *e_new = (uint8_t*) calloc((size_t) 2 * nq + 4, 1), *s_new = (uint8_t*) calloc((size_t) 2 * nq + 4, 1);
	for (int y = 0; y < h; y += 2) {
		const int first = y == 0;
		const uint16_t *tbl = first ? HT_VLC_DEC0 : HT_VLC_DEC1;
		memset(e_new, 0, (size_t) 2 * nq + 4);
		memset(s_new, 0, (size_t) 2 * nq + 4);
		int prev_rho = 0;
		for (int qx = 0; qx < nq; qx += 2) {
			uint32_t info[2] = {0, 0};
			int U[2] = {0, 0};
			const int present[2] = {1, qx + 1 < nq};
			for (int k = 0; k < 2; ++k) {
				if (!present[k]) break;
				const int q = qx + k;
				int cq;
				if (first) cq = (prev_rho >> 1) | (prev_rho & 1);
				else {
					const int snw = q ? s_bot[2 * q - 1] : 0, sn = s_bot[2 * q], sne = s_bot[2 * q + 1], snf = s_bot[2 * q + 2];
					cq = (snw | sn) | (((prev_rho >> 2) | (prev_rho >> 3)) & 1) << 1 | (sne | snf) << 2;
				}
				const uint32_t v = rev_peek(&vlc);
				uint32_t t = tbl[(cq << 7) | (v & 0x7F)];
				if (cq == 0 && !mel_event_read(&mel)) t = 0; /* an all-zero quad in the zero context costs no VLC bits */
				rev_skip(&vlc, (int) (t & 7));
				info[k] = t;
				prev_rho = (int) (t >> 4) & 15;
			}
			/* U-VLC of the pair */
			{
				uint32_t v = rev_peek(&vlc);
				int used = 0;
				const int uo0 = (int) (info[0] >> 3) & 1, uo1 = (int) (info[1] >> 3) & 1;
				int u0 = 0, u1 = 0;
				if (first && uo0 && uo1) {
					if (mel_event_read(&mel)) { /* both u exceed 2: coded as u - 2 */
						const int p0 = uvlc_prefix(&v, &used), p1 = uvlc_prefix(&v, &used);
						u0 = uvlc_suffix(p0, &v, &used) + 2;
						u1 = uvlc_suffix(p1, &v, &used) + 2;
					} else {
						const int p0 = uvlc_prefix(&v, &used);
						if (p0 > 2) { /* the second quad's u is 1 or 2: a single bit */
							u1 = (int) (v & 1) + 1; v >>= 1; used++;
							u0 = uvlc_suffix(p0, &v, &used);
						} else {
							const int p1 = uvlc_prefix(&v, &used);
							u0 = uvlc_suffix(p0, &v, &used);
							u1 = uvlc_suffix(p1, &v, &used);
						}
					}
				} else {
					const int p0 = uo0 ? uvlc_prefix(&v, &used) : 0, p1 = uo1 ? uvlc_prefix(&v, &used) : 0;
					if (uo0) u0 = uvlc_suffix(p0, &v, &used);
					if (uo1) u1 = uvlc_suffix(p1, &v, &used);
				}
				rev_skip(&vlc, used);
				U[0] = u0; U[1] = u1;
			}
			for (int k = 0; k < 2; ++k) {
				if (!present[k]) break;
				const int q = qx + k;
				const int rho = (int) (info[k] >> 4) & 15, ek = (int) (info[k] >> 12) & 15, e1 = (int) (info[k] >> 8) & 15;
				int kappa = 1;
				if (!first && (rho & (rho - 1))) {
					int emax = q ? e_bot[2 * q - 1] : 0;
					if (e_bot[2 * q] > emax) emax = e_bot[2 * q];
					if (e_bot[2 * q + 1] > emax) emax = e_bot[2 * q + 1];
					if (e_bot[2 * q + 2] > emax) emax = e_bot[2 * q + 2];
					kappa = emax - 1 > 1 ? emax - 1 : 1;
				}
				const int Uq = U[k] + kappa;
				for (int i = 0; i < 4; ++i) {
					const int xx = 2 * q + (i >> 1), yy = y + (i & 1);
					int32_t val = 0;
					int e = 0;
					if (rho >> i & 1) {
						int m = Uq - (ek >> i & 1);
						if (m > 31) m = 31; /* only a corrupt stream gets here; keeps the shifts defined */
						const uint32_t b = fwd_peek(&ms);
						fwd_skip(&ms, m);
						uint32_t vn = b & ((1u << m) - 1u);
						vn |= (uint32_t) (e1 >> i & 1) << m; /* the implicit top bit of a sample that reaches the exponent bound */
						vn |= 1;                              /* reconstruct at the centre of the bin */
						val = (int32_t) ((b << 31) | ((vn + 2) << (p - 1)));
						e = bits_of(vn);
					}
					if (xx < w && yy < h) out[(size_t) yy * stride + xx] = val;
					if (i & 1) { e_new[xx] = (uint8_t) e; s_new[xx] = (uint8_t) (rho >> i & 1); }
				}
			}
		}
		uint8_t *t = e_bot; e_bot = e_new; e_new = t;
		t = s_bot; s_bot = s_new; s_new = t;
	}
	free(e_bot); free(s_bot); free(e_new); free(s_new);
	return 0;
}
