// K4' / K5': the HTJ2K block coder (Rec. ITU-T T.814 | ISO/IEC 15444-15), cleanup pass -- what the reference runs when
// the code-block style has the HT bit (grk_compress -M 64): its encoder emits exactly one cleanup pass per block.
//
//   T1HT::preEncode        t1/t1_ht/T1HT.cpp:56-103               sign-magnitude, MSB aligned (fused into the load here)
//   ojph_encode_codeblock  t1_ht/coding/ojph_block_encoder.cpp:465-938   MagSgn + MEL + VLC byte streams
//   ojph_decode_codeblock  t1_ht/coding/ojph_block_decoder.cpp:687-1200  cleanup pass
//   T1HT::postDecode       t1/t1_ht/T1HT.cpp:176-251              de-quantisation (fused into the store here)
//
// One THREAD per code block.  The HT coder has no adaptive arithmetic coder: a quad (2x2 samples) costs one table look-up,
// a few exponent comparisons and three bit-stream appends, an order of magnitude less work per sample than the MQ path, but
// its three byte streams are bit-stuffed (the byte after 0xFF / after a byte > 0x8F carries seven bits), so packing is
// sequential within a block and the parallelism is across blocks, as in t1_mq_kernel.  The line state a row of quads hands to
// the next one (exponent and significance of its bottom samples) lives in shared memory, two byte rows per thread.
// MagSgn grows from the front of the block's scratch area, VLC from its end, MEL sits in between; the thread that coded the
// block moves MEL and VLC up behind MagSgn and patches the 12-bit suffix length, so the compaction kernel of the MQ path
// (t1_offsets_kernel / t1_gather_kernel) serves both coders.
#include "common.cuh"
#include <mutex>

namespace gb {

#include "ht_tables.inc"

constexpr int HT_THREADS = 64;      // code blocks per CTA
constexpr int HT_LINE = 2 * 32 + 4; // bottom samples of a row of quads (blocks up to 64 wide) + the two neighbours past the ends
constexpr int HT_MEL_CAP = 1536, HT_VLC_CAP = 2560;

// the derived CxtVLC tables live in global memory and are read through the read-only path: the lanes of a warp index them with
// unrelated values, which constant memory would serialise
__device__ uint16_t c_ht_enc0[2048], c_ht_enc1[2048], c_ht_dec0[1024], c_ht_dec1[1024];
__device__ const int c_mel_e[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5};

__device__ __forceinline__ int ht_bits(uint32_t v) { return 32 - __clz(v); }

// ---- writers -----------------------------------------------------------------------------------------------------------
struct HtFwd { uint8_t *buf; int pos, cap, used, limit; uint32_t acc; };   // MagSgn: LSB first

__device__ __forceinline__ void ht_ms_put(HtFwd &s, uint32_t bits, int n) {
	while (n > 0) {
		const int take = min(s.limit - s.used, n);
		s.acc |= (bits & ((1u << take) - 1u)) << s.used;
		s.used += take;
		bits >>= take;
		n -= take;
		if (s.used == s.limit) {
			if (s.pos < s.cap) s.buf[s.pos] = (uint8_t) s.acc;
			s.pos++;
			s.limit = s.acc == 0xFFu ? 7 : 8;
			s.acc = 0;
			s.used = 0;
		}
	}
}

struct HtMel { uint8_t *buf; int pos, cap, left, run, k, threshold; uint32_t acc; }; // MSB first

__device__ __forceinline__ void ht_mel_bit(HtMel &m, int v) {
	m.acc = (m.acc << 1) | (uint32_t) v;
	if (--m.left == 0) {
		if (m.pos < m.cap) m.buf[m.pos] = (uint8_t) m.acc;
		m.pos++;
		m.left = m.acc == 0xFFu ? 7 : 8;
		m.acc = 0;
	}
}
__device__ __forceinline__ void ht_mel_event(HtMel &m, int one) {
	if (!one) {
		if (++m.run >= m.threshold) {
			ht_mel_bit(m, 1);
			m.run = 0;
			m.k = min(m.k + 1, 12);
			m.threshold = 1 << c_mel_e[m.k];
		}
	} else {
		ht_mel_bit(m, 0);
		for (int t = c_mel_e[m.k]; t > 0;) ht_mel_bit(m, (m.run >> --t) & 1);
		m.run = 0;
		m.k = max(m.k - 1, 0);
		m.threshold = 1 << c_mel_e[m.k];
	}
}

struct HtRev { uint8_t *end; int pos, cap, used, prev_gt_8f; uint32_t acc; }; // VLC: downwards, LSB first

__device__ __forceinline__ void ht_vlc_put(HtRev &s, uint32_t bits, int n) {
	while (n > 0) {
		int room = 8 - s.prev_gt_8f - s.used;
		const int take = min(room, n);
		s.acc |= (bits & ((1u << take) - 1u)) << s.used;
		s.used += take;
		room -= take;
		n -= take;
		bits >>= take;
		if (room == 0) {
			if (s.prev_gt_8f && s.acc != 0x7Fu) { s.prev_gt_8f = 0; continue; } // the eighth bit is usable after all
			if (s.pos < s.cap) s.end[-s.pos] = (uint8_t) s.acc;
			s.pos++;
			s.prev_gt_8f = s.acc > 0x8Fu;
			s.acc = 0;
			s.used = 0;
		}
	}
}

// U-VLC code of u: prefix 1 / 01 / 001 / 000 (LSB first), then 0, 0, 1 or 5 suffix bits.  Packed: pre | pre_len << 8 | suf << 12 | suf_len << 20
__device__ __forceinline__ uint32_t ht_uvlc(int u) {
	if (u == 0) return 0;
	if (u == 1) return 1u | 1u << 8;
	if (u == 2) return 2u | 2u << 8;
	if (u <= 4) return 4u | 3u << 8 | (uint32_t) (u - 3) << 12 | 1u << 20;
	return 0u | 3u << 8 | (uint32_t) (u - 5) << 12 | 5u << 20;
}

// quantised sign-magnitude sample (T1HT.cpp:68-100): reversible |x| << shift; irreversible (int) (x * (1 / stepsize) * 2^shift)
__device__ __forceinline__ uint32_t ht_sample(const EncBlock &B, int x, int y, int shift, float inv) {
	const int32_t t = B.src[(size_t) y * B.stride + x];
	if (B.reversible) return (t >= 0 ? 0u : 0x80000000u) | ((uint32_t) abs(t) << shift);
	const int32_t q = (int32_t) __fmul_rn(__fmul_rn((float) t, inv), (float) (1 << shift)); // truncation, as the C cast
	return (q >= 0 ? 0u : 0x80000000u) | (uint32_t) abs(q);
}

__global__ void __launch_bounds__(HT_THREADS) t1_ht_encode_kernel(const EncBlock *__restrict__ blocks, uint32_t nblocks,
		uint8_t *__restrict__ scratch, EncResult *__restrict__ results, uint32_t *__restrict__ rates, double *__restrict__ dists) {
	__shared__ uint8_t line[HT_THREADS][4][HT_LINE]; // [0] / [1]: exponents, [2] / [3]: significance; two generations each
	const uint32_t bid = blockIdx.x * HT_THREADS + threadIdx.x;
	if (bid >= nblocks) return;
	const EncBlock B = blocks[bid];
	const int w = B.w, h = B.h;
	EncResult res = {1u, 1u, 0u, 0u, 0ull}; // T1HT::encode: always one pass, numbps = 1 (T1HT.cpp:125-128)
	if (w == 0 || h == 0) { // a zero-area block of an empty band: the reference never reaches the coder with it
		res.numbps = 0; res.numpasses = 0;
		results[bid] = res;
		return;
	}
	const int missing = B.band_numbps; // k_msbs = band->numbps - cblk->numbps with a fresh block (Tier1.cpp:86)
	const int p = 30 - missing;
	const int shift = B.reversible ? 31 - (missing + 1) : 31 - (missing + 1) - 11;
	const float inv = __fdiv_rn(1.0f, B.stepsize); // Tier1.cpp:78
	uint8_t *out = scratch + B.scratch_off + 1;
	const int cap = (int) B.scratch_cap - 1;
	HtFwd ms = {out, 0, cap - HT_MEL_CAP - HT_VLC_CAP, 0, 8, 0u};
	HtMel mel = {out + ms.cap, 0, HT_MEL_CAP, 8, 0, 0, 1, 0u};
	HtRev vlc = {out + cap - 1, 1, HT_VLC_CAP, 4, 1, 0xFu};
	vlc.end[0] = 0xFF;
	uint8_t *eb = line[threadIdx.x][0], *en = line[threadIdx.x][1], *sb = line[threadIdx.x][2], *sn = line[threadIdx.x][3];
	for (int i = 0; i < HT_LINE; ++i) { eb[i] = 0; sb[i] = 0; en[i] = 0; sn[i] = 0; }
	// line arrays are indexed by column + 1, so that the north-west neighbour of the first quad reads a zero
	const int nq = (w + 1) >> 1;
	uint32_t nquads = 0;
	for (int y = 0; y < h; y += 2) {
		const bool first = y == 0;
		const uint16_t *tbl = first ? c_ht_enc0 : c_ht_enc1;
		int prev_rho = 0;
		for (int qx = 0; qx < nq; qx += 2) { // quads are coded in pairs
			int u[2] = {0, 0};
			#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const int q = qx + k;
				if (q >= nq) break;
				int rho = 0, emax = 0, e[4];
				uint32_t v[4];
				#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int xx = 2 * q + (i >> 1), yy = y + (i & 1);
					e[i] = 0; v[i] = 0;
					if (xx < w && yy < h) {
						const uint32_t t = ht_sample(B, xx, yy, shift, inv);
						uint32_t val = ((t + t) >> p) & ~1u; // 2 * mu_p
						if (val) {
							rho |= 1 << i;
							e[i] = ht_bits(val - 1);
							emax = max(emax, e[i]);
							v[i] = (val - 2) + (t >> 31); // 2 (mu_p - 1) + sign
						}
					}
				}
				int cq, kappa = 1;
				if (first) cq = (prev_rho >> 1) | (prev_rho & 1);
				else {
					const int c = 2 * q + 1; // column + 1
					cq = (sb[c - 1] | sb[c]) | (((prev_rho >> 2) | (prev_rho >> 3)) & 1) << 1 | (sb[c + 1] | sb[c + 2]) << 2;
					if (rho & (rho - 1)) kappa = max(max(max((int) eb[c - 1], (int) eb[c]), max((int) eb[c + 1], (int) eb[c + 2])) - 1, 1);
				}
				const int U = max(emax, kappa);
				u[k] = U - kappa;
				int eps = 0;
				if (u[k] > 0) {
					#pragma unroll
					for (int i = 0; i < 4; ++i) eps |= (e[i] == emax) << i;
				}
				const uint32_t tuple = __ldg(tbl + ((cq << 8) | (rho << 4) | eps));
				ht_vlc_put(vlc, tuple >> 8, (tuple >> 4) & 7);
				if (cq == 0) ht_mel_event(mel, rho != 0);
				#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int m = (rho >> i & 1) ? U - (int) (tuple >> i & 1) : 0;
					ht_ms_put(ms, v[i] & ((1u << m) - 1u), m);
				}
				en[2 * q + 1] = (uint8_t) e[1]; en[2 * q + 2] = (uint8_t) e[3];
				sn[2 * q + 1] = (uint8_t) (rho >> 1 & 1); sn[2 * q + 2] = (uint8_t) (rho >> 3 & 1);
				prev_rho = rho;
				nquads++;
			}
			// the U-VLC codes of the pair: both prefixes, then both suffixes (first row: T.814 7.3.6 special cases)
			uint32_t c0, c1;
			if (first && u[0] > 0 && u[1] > 0) ht_mel_event(mel, min(u[0], u[1]) > 2);
			if (first && u[0] > 2 && u[1] > 2) { c0 = ht_uvlc(u[0] - 2); c1 = ht_uvlc(u[1] - 2); }
			else if (first && u[0] > 2 && u[1] > 0) { c0 = ht_uvlc(u[0]); c1 = (uint32_t) (u[1] - 1) | 1u << 8; }
			else { c0 = ht_uvlc(u[0]); c1 = ht_uvlc(u[1]); }
			ht_vlc_put(vlc, c0 & 0xFF, (c0 >> 8) & 0xF);
			ht_vlc_put(vlc, c1 & 0xFF, (c1 >> 8) & 0xF);
			ht_vlc_put(vlc, (c0 >> 12) & 0xFF, (c0 >> 20) & 0xF);
			ht_vlc_put(vlc, (c1 >> 12) & 0xFF, (c1 >> 20) & 0xF);
		}
		uint8_t *t = eb; eb = en; en = t;
		t = sb; sb = sn; sn = t;
		for (int i = 0; i < HT_LINE; ++i) { en[i] = 0; sn[i] = 0; }
	}
	// ---- termination: the open MEL and VLC bytes share one byte when their used bits do not collide ----
	if (mel.run > 0) ht_mel_bit(mel, 1);
	{
		const uint32_t mel_tmp = (mel.acc << mel.left) & 0xFFu;
		const uint32_t mel_mask = (0xFFu << mel.left) & 0xFFu, vlc_mask = 0xFFu >> (8 - vlc.used);
		if ((mel_mask | vlc_mask) != 0) {
			const uint32_t fuse = mel_tmp | vlc.acc;
			if ((((fuse ^ mel_tmp) & mel_mask) | ((fuse ^ vlc.acc) & vlc_mask)) == 0 && fuse != 0xFFu && vlc.pos > 1) {
				if (mel.pos < mel.cap) mel.buf[mel.pos] = (uint8_t) fuse;
				mel.pos++;
			} else {
				if (mel.pos < mel.cap) mel.buf[mel.pos] = (uint8_t) mel_tmp;
				mel.pos++;
				if (vlc.pos < vlc.cap) vlc.end[-vlc.pos] = (uint8_t) vlc.acc;
				vlc.pos++;
			}
		}
	}
	if (ms.used) { // pad the open MagSgn byte with ones; a padded 0xFF is dropped
		ms.acc |= ((1u << (ms.limit - ms.used)) - 1u) << ms.used;
		if (ms.acc != 0xFFu) {
			if (ms.pos < ms.cap) ms.buf[ms.pos] = (uint8_t) ms.acc;
			ms.pos++;
		}
	} else if (ms.limit == 7) ms.pos--;
	const bool overflow = ms.pos > ms.cap || mel.pos > mel.cap || vlc.pos > vlc.cap;
	const int total = ms.pos + mel.pos + vlc.pos;
	if (!overflow) {
		// MEL and VLC move up behind MagSgn (both lie above their destination, so an ascending copy is safe)
		for (int i = 0; i < mel.pos; ++i) out[ms.pos + i] = mel.buf[i];
		const uint8_t *vsrc = vlc.end - vlc.pos + 1;
		for (int i = 0; i < vlc.pos; ++i) out[ms.pos + mel.pos + i] = vsrc[i];
		const int scup = mel.pos + vlc.pos; // the last twelve bits of the block locate the MEL + VLC suffix
		out[total - 1] = (uint8_t) (scup >> 4);
		out[total - 2] = (uint8_t) ((out[total - 2] & 0xF0) | (scup & 0xF));
	}
	res.numpasses = overflow ? 0xFFFFFFFFu : 1u;
	res.data_len = overflow ? 0u : (uint32_t) total;
	res.decisions = nquads;
	if (B.max_passes) { rates[B.pass_offset] = res.data_len; dists[B.pass_offset] = 0.0; }
	results[bid] = res;
}

// ---- readers -----------------------------------------------------------------------------------------------------------
struct HtFwdR { const uint8_t *p; int size, pos, bits, unstuff; uint64_t acc; };

__device__ __forceinline__ uint32_t ht_fwd_peek(HtFwdR &r) {
	while (r.bits <= 32) { // bytes past the end read as 0xFF; after a 0xFF the next byte gives 7 bits
		const uint32_t d = r.pos < r.size ? r.p[r.pos] : 0xFFu;
		r.pos++;
		r.acc |= (uint64_t) d << r.bits;
		r.bits += 8 - r.unstuff;
		r.unstuff = d == 0xFFu;
	}
	return (uint32_t) r.acc;
}

struct HtMelR { const uint8_t *p; int size, pos, bits, unstuff, k, run, one; uint32_t acc; };

__device__ __forceinline__ int ht_mel_next_bit(HtMelR &m) {
	if (m.bits == 0) {
		uint32_t d = m.pos < m.size ? m.p[m.pos] : 0xFFu;
		if (m.pos == m.size - 1) d |= 0xFu; // the last byte of the segment shares its low nibble with the suffix length
		m.pos++;
		const int n = 8 - m.unstuff;
		m.acc = d & ((1u << n) - 1u);
		m.bits = n;
		m.unstuff = d == 0xFFu;
	}
	m.bits--;
	return (int) (m.acc >> m.bits) & 1;
}
__device__ __forceinline__ int ht_mel_event_read(HtMelR &m) {
	if (m.run == 0 && !m.one) {
		const int e = c_mel_e[m.k];
		if (ht_mel_next_bit(m)) { m.run = 1 << e; m.one = 0; m.k = min(m.k + 1, 12); }
		else {
			int r = 0;
			for (int i = 0; i < e; ++i) r = (r << 1) | ht_mel_next_bit(m);
			m.run = r; m.one = 1;
			m.k = max(m.k - 1, 0);
		}
	}
	if (m.run > 0) { m.run--; return 0; }
	m.one = 0;
	return 1;
}

struct HtRevR { const uint8_t *base; int pos, bits, unstuff; uint64_t acc; };

__device__ __forceinline__ uint32_t ht_rev_peek(HtRevR &r) {
	while (r.bits <= 32) {
		const uint32_t d = r.pos >= 0 ? r.base[r.pos] : 0u;
		r.pos--;
		const int n = 8 - ((r.unstuff && (d & 0x7Fu) == 0x7Fu) ? 1 : 0);
		r.acc |= (uint64_t) d << r.bits;
		r.bits += n;
		r.unstuff = d > 0x8Fu;
	}
	return (uint32_t) r.acc;
}

__device__ __forceinline__ int ht_uvlc_prefix(uint32_t &v, int &used) {
	int pv, pl;
	if (v & 1) { pv = 1; pl = 1; } else if (v & 2) { pv = 2; pl = 2; } else if (v & 4) { pv = 3; pl = 3; } else { pv = 5; pl = 3; }
	v >>= pl; used += pl;
	return pv;
}
__device__ __forceinline__ int ht_uvlc_suffix(int prefix, uint32_t &v, int &used) {
	const int sl = prefix == 3 ? 1 : prefix == 5 ? 5 : 0;
	const int s = (int) (v & ((1u << sl) - 1u));
	v >>= sl; used += sl;
	return prefix + s;
}

__global__ void __launch_bounds__(HT_THREADS) t1_ht_decode_kernel(const DecBlock *__restrict__ blocks, const DecInput *__restrict__ inputs,
		uint32_t nblocks, const uint8_t *__restrict__ data) {
	__shared__ uint8_t line[HT_THREADS][4][HT_LINE];
	const uint32_t bid = blockIdx.x * HT_THREADS + threadIdx.x;
	if (bid >= nblocks) return;
	const DecBlock B = blocks[bid];
	const DecInput I = inputs[bid];
	const int w = B.w, h = B.h;
	if (w == 0 || h == 0) return;
	const int lcup = (int) I.data_len;
	const uint8_t *D = data + I.data_offset;
	int scup = 0;
	bool ok = I.numpasses != 0 && lcup >= 2;
	if (ok) {
		scup = ((int) D[lcup - 1] << 4) + (D[lcup - 2] & 0xF);
		ok = scup <= lcup && scup >= 2;
	}
	if (!ok) { // no data for this block (or an inconsistent suffix length: T1HT::decode leaves the block undecoded): zeros
		for (int y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) B.dst[(size_t) y * B.stride + x] = 0;
		return;
	}
	const int missing = (int) B.band_numbps - (int) I.numbps; // k_msbs (Tier1.cpp:166)
	const int p = 30 - missing;
	const int dshift = 31 - (missing + 1); // T1HT.cpp:213
	HtFwdR ms = {D, lcup - scup, 0, 0, 0, 0ull};
	HtMelR mel = {D + lcup - scup, scup - 1, 0, 0, 0, 0, 0, 0, 0u};
	HtRevR vlc = {D, lcup - 3, 0, 0, 0ull};
	{
		const uint32_t d = D[lcup - 2]; // its upper nibble opens the VLC stream
		vlc.acc = d >> 4;
		vlc.bits = 4 - ((vlc.acc & 7) == 7);
		vlc.unstuff = (d | 0xF) > 0x8F;
	}
	uint8_t *eb = line[threadIdx.x][0], *en = line[threadIdx.x][1], *sb = line[threadIdx.x][2], *sn = line[threadIdx.x][3];
	for (int i = 0; i < HT_LINE; ++i) { eb[i] = 0; sb[i] = 0; en[i] = 0; sn[i] = 0; }
	const int nq = (w + 1) >> 1;
	for (int y = 0; y < h; y += 2) {
		const bool first = y == 0;
		const uint16_t *tbl = first ? c_ht_dec0 : c_ht_dec1;
		int prev_rho = 0;
		for (int qx = 0; qx < nq; qx += 2) {
			uint32_t info[2] = {0, 0};
			#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const int q = qx + k;
				if (q >= nq) break;
				int cq;
				if (first) cq = (prev_rho >> 1) | (prev_rho & 1);
				else {
					const int c = 2 * q + 1;
					cq = (sb[c - 1] | sb[c]) | (((prev_rho >> 2) | (prev_rho >> 3)) & 1) << 1 | (sb[c + 1] | sb[c + 2]) << 2;
				}
				const uint32_t v = ht_rev_peek(vlc);
				uint32_t t = __ldg(tbl + ((cq << 7) | (v & 0x7F)));
				if (cq == 0 && !ht_mel_event_read(mel)) t = 0; // an all-zero quad in the zero context costs no VLC bits
				vlc.acc >>= (t & 7); vlc.bits -= (int) (t & 7);
				info[k] = t;
				prev_rho = (int) (t >> 4) & 15;
			}
			int u0 = 0, u1 = 0;
			{
				uint32_t v = ht_rev_peek(vlc);
				int used = 0;
				const int uo0 = (int) (info[0] >> 3) & 1, uo1 = (int) (info[1] >> 3) & 1;
				if (first && uo0 && uo1) {
					if (ht_mel_event_read(mel)) { // both u exceed 2: coded as u - 2
						const int p0 = ht_uvlc_prefix(v, used), p1 = ht_uvlc_prefix(v, used);
						u0 = ht_uvlc_suffix(p0, v, used) + 2;
						u1 = ht_uvlc_suffix(p1, v, used) + 2;
					} else {
						const int p0 = ht_uvlc_prefix(v, used);
						if (p0 > 2) { // the second quad's u is 1 or 2: a single bit
							u1 = (int) (v & 1) + 1; v >>= 1; used++;
							u0 = ht_uvlc_suffix(p0, v, used);
						} else {
							const int p1 = ht_uvlc_prefix(v, used);
							u0 = ht_uvlc_suffix(p0, v, used);
							u1 = ht_uvlc_suffix(p1, v, used);
						}
					}
				} else {
					const int p0 = uo0 ? ht_uvlc_prefix(v, used) : 0, p1 = uo1 ? ht_uvlc_prefix(v, used) : 0;
					if (uo0) u0 = ht_uvlc_suffix(p0, v, used);
					if (uo1) u1 = ht_uvlc_suffix(p1, v, used);
				}
				vlc.acc >>= used; vlc.bits -= used;
			}
			#pragma unroll
			for (int k = 0; k < 2; ++k) {
				const int q = qx + k;
				if (q >= nq) break;
				const int rho = (int) (info[k] >> 4) & 15, ek = (int) (info[k] >> 12) & 15, e1 = (int) (info[k] >> 8) & 15;
				int kappa = 1;
				const int c = 2 * q + 1;
				if (!first && (rho & (rho - 1))) kappa = max(max(max((int) eb[c - 1], (int) eb[c]), max((int) eb[c + 1], (int) eb[c + 2])) - 1, 1);
				const int Uq = (k ? u1 : u0) + kappa;
				#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int xx = 2 * q + (i >> 1), yy = y + (i & 1);
					uint32_t val = 0;
					int e = 0;
					if (rho >> i & 1) {
						const int m = Uq - (ek >> i & 1);
						const uint32_t b = ht_fwd_peek(ms);
						ms.acc >>= m; ms.bits -= m;
						uint32_t vn = b & ((1u << m) - 1u);
						vn |= (uint32_t) (e1 >> i & 1) << m; // the implicit top bit of a sample that reaches the exponent bound
						vn |= 1;                              // reconstruct at the centre of the bin
						val = (b << 31) | ((vn + 2) << (p - 1));
						e = ht_bits(vn);
					}
					if (xx < w && yy < h) { // T1HT::postDecode, T1HT.cpp:211-236
						const int32_t mag = (int32_t) (val & 0x7FFFFFFFu);
						int32_t o;
						if (B.reversible) { const int32_t r = mag >> dshift; o = (val >> 31) ? -r : r; }
						else { const float f = __fmul_rn((float) mag, B.stepsize); o = __float_as_int((val >> 31) ? -f : f); }
						B.dst[(size_t) yy * B.stride + xx] = o;
					}
					if (i & 1) { en[xx + 1] = (uint8_t) e; sn[xx + 1] = (uint8_t) (rho >> i & 1); }
				}
			}
		}
		uint8_t *t = eb; eb = en; en = t;
		t = sb; sb = sn; sn = t;
		for (int i = 0; i < HT_LINE; ++i) { en[i] = 0; sn[i] = 0; }
	}
}

// the four derived CxtVLC tables, once per device
static void ensure_ht_tables() {
	static std::mutex mu;
	static uint64_t ready[4] = {0, 0, 0, 0};
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(mu);
	if (dev >= 0 && dev < 256 && (ready[dev >> 6] >> (dev & 63) & 1)) return;
	cudaMemcpyToSymbol(c_ht_enc0, HT_VLC_ENC0, sizeof(HT_VLC_ENC0));
	cudaMemcpyToSymbol(c_ht_enc1, HT_VLC_ENC1, sizeof(HT_VLC_ENC1));
	cudaMemcpyToSymbol(c_ht_dec0, HT_VLC_DEC0, sizeof(HT_VLC_DEC0));
	cudaMemcpyToSymbol(c_ht_dec1, HT_VLC_DEC1, sizeof(HT_VLC_DEC1));
	if (dev >= 0 && dev < 256) ready[dev >> 6] |= 1ull << (dev & 63);
}

uint32_t t1_ht_scratch_extra() { return HT_MEL_CAP + HT_VLC_CAP + 64; }

void launch_t1_ht_encode(const EncBlock *blocks, uint32_t nblocks, uint8_t *scratch, EncResult *results, uint32_t *rates, double *dists,
		cudaStream_t s) {
	if (!nblocks) return;
	ensure_ht_tables();
	t1_ht_encode_kernel<<<(nblocks + HT_THREADS - 1) / HT_THREADS, HT_THREADS, 0, s>>>(blocks, nblocks, scratch, results, rates, dists);
}

void launch_t1_ht_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data, cudaStream_t s) {
	if (!nblocks) return;
	ensure_ht_tables();
	t1_ht_decode_kernel<<<(nblocks + HT_THREADS - 1) / HT_THREADS, HT_THREADS, 0, s>>>(blocks, inputs, nblocks, data);
}

} // namespace gb
