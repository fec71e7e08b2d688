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
#include <cstdlib>

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

// ---- encoder, one WARP per code block ---------------------------------------------------------------------------------
// A row of quads of a block (at most 32 quads: blocks are at most 64 wide) is one quad per lane.  What a quad sends depends on
// its left neighbour's significance pattern and on the bottom samples of the quads above: lane-to-lane shuffles and four
// registers of line state per lane, so the whole analysis of a row -- quantisation, exponents, context, kappa, CxtVLC look-up,
// U-VLC of the pair, MagSgn field widths -- runs on all lanes at once.  The lanes then lay their bits side by side in two
// per-warp bit buffers in shared memory (prefix sum of the lengths, atomicOr of the pieces); only the bit-stuffed byte packing of
// the three streams, which is inherently sequential (the byte after 0xFF / after a byte > 0x8F holds seven bits), and the MEL
// state machine (adaptive, but at most three events per quad pair) are left to lane 0.
constexpr int HTW_WARPS = 4;          // code blocks per CTA
constexpr int HTW_MS_WORDS = 128;     // MagSgn bits of one row of quads: 32 quads x 4 samples x at most 31 bits
constexpr int HTW_VLC_WORDS = 18;     // VLC bits of one row: 16 pairs x at most 30 bits

__device__ __forceinline__ uint32_t ht_take_bits(const uint32_t *buf, int pos, int n) { // n <= 8 bits at bit position pos
	const int wi = pos >> 5, sh = pos & 31;
	uint32_t v = buf[wi] >> sh;
	if (sh + n > 32) v |= buf[wi + 1] << (32 - sh);
	return v & ((1u << n) - 1u);
}
__device__ __forceinline__ void ht_or_bits(uint32_t *buf, int pos, uint32_t bits, int n) { // n <= 31 bits, shared memory
	if (n <= 0) return;
	const int wi = pos >> 5, sh = pos & 31;
	atomicOr(buf + wi, bits << sh);
	if (sh + n > 32) atomicOr(buf + wi + 1, bits >> (32 - sh));
}

__global__ void __launch_bounds__(HTW_WARPS * 32) t1_ht_encode_kernel(const EncBlock *__restrict__ blocks, uint32_t nblocks,
		uint8_t *__restrict__ scratch, EncResult *__restrict__ results, uint32_t *__restrict__ rates, double *__restrict__ dists) {
	__shared__ uint32_t ms_bits[HTW_WARPS][HTW_MS_WORDS + 2], vlc_bits[HTW_WARPS][HTW_VLC_WORDS + 2];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bid = blockIdx.x * HTW_WARPS + wid;
	if (bid >= nblocks) return;
	const EncBlock B = blocks[bid];
	const int w = B.w, h = B.h;
	if (w == 0 || h == 0) { // a zero-area block of an empty band: the reference never reaches the coder with it
		if (lane == 0) { EncResult r = {0u, 0u, 0u, 0u, 0ull}; results[bid] = r; }
		return;
	}
	const int missing = B.band_numbps; // k_msbs = band->numbps - cblk->numbps with a fresh block (Tier1.cpp:86)
	const int p = 30 - missing;
	const int shift = B.reversible ? 31 - (missing + 1) : 31 - (missing + 1) - 11;
	const float inv = __fdiv_rn(1.0f, B.stepsize); // Tier1.cpp:78
	uint8_t *out = scratch + B.scratch_off + 1;
	const int cap = (int) B.scratch_cap - 1;
	// the MEL writer lives in lane 0's registers; the MagSgn and VLC writers are warp-uniform state (every lane holds the same
	// values): position, open bits (count, value) and -- MagSgn -- whether the last byte written was 0xFF / -- VLC -- that byte
	HtMel mel = {out + (cap - HT_MEL_CAP - HT_VLC_CAP), 0, HT_MEL_CAP, 8, 0, 0, 1, 0u};
	uint8_t *const ms_buf = out;
	const int ms_cap = cap - HT_MEL_CAP - HT_VLC_CAP;
	int ms_pos = 0, ms_used = 0, ms_limit = 8;
	uint32_t ms_acc = 0;
	uint8_t *const vlc_end = out + cap - 1;
	int vlc_pos = 1, vlc_used = 4;
	uint32_t vlc_acc = 0xFu, vlc_prev = 0xFFu; // the first byte starts with four bits set, "after" a byte of 0xFF (vlc_init)
	if (lane == 0) vlc_end[0] = 0xFF;
	uint32_t *msb = ms_bits[wid], *vlb = vlc_bits[wid];
	for (int i = lane; i < HTW_MS_WORDS + 2; i += 32) msb[i] = 0;
	for (int i = lane; i < HTW_VLC_WORDS + 2; i += 32) vlb[i] = 0;
	__syncwarp();
	const int nq = (w + 1) >> 1;
	const bool active = lane < nq;
	int pe_bl = 0, pe_br = 0, ps_bl = 0, ps_br = 0; // line state: exponent / significance of the bottom samples of the quad above
	for (int y = 0; y < h; y += 2) {
		const bool first = y == 0;
		// ---- the quad of this lane -------------------------------------------------------------------------------------
		int rho = 0, emax = 0, e[4];
		uint32_t v[4];
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int xx = 2 * lane + (i >> 1), yy = y + (i & 1);
			e[i] = 0; v[i] = 0;
			if (active && xx < w && yy < h) {
				const uint32_t t = ht_sample(B, xx, yy, shift, inv);
				const uint32_t val = ((t + t) >> p) & ~1u; // 2 * mu_p
				if (val) {
					rho |= 1 << i;
					e[i] = ht_bits(val - 1);
					emax = max(emax, e[i]);
					v[i] = (val - 2) + (t >> 31); // 2 (mu_p - 1) + sign
				}
			}
		}
		int rho_l = __shfl_up_sync(0xffffffffu, rho, 1);
		if (lane == 0) rho_l = 0;
		int cq, kappa = 1;
		{
			int s_nw = __shfl_up_sync(0xffffffffu, ps_br, 1), e_nw = __shfl_up_sync(0xffffffffu, pe_br, 1);
			const int s_nf = __shfl_down_sync(0xffffffffu, ps_bl, 1), e_nf = __shfl_down_sync(0xffffffffu, pe_bl, 1); // lanes past the block hold zeros
			if (lane == 0) { s_nw = 0; e_nw = 0; }
			if (first) cq = (rho_l >> 1) | (rho_l & 1);
			else {
				cq = (s_nw | ps_bl) | (((rho_l >> 2) | (rho_l >> 3)) & 1) << 1 | (ps_br | (lane == 31 ? 0 : s_nf)) << 2;
				if (rho & (rho - 1)) kappa = max(max(max(e_nw, pe_bl), max(pe_br, lane == 31 ? 0 : e_nf)) - 1, 1);
			}
		}
		const int U = max(emax, kappa);
		const int u = active ? U - kappa : 0;
		int eps = 0;
		if (u > 0) {
			#pragma unroll
			for (int i = 0; i < 4; ++i) eps |= (e[i] == emax) << i;
		}
		const uint32_t tuple = active ? __ldg((first ? c_ht_enc0 : c_ht_enc1) + ((cq << 8) | (rho << 4) | eps)) : 0u;
		// ---- MEL events of the row, in coding order: quad, quad, pair ----------------------------------------------------
		const int u_r = __shfl_down_sync(0xffffffffu, u, 1); // the partner quad's u (0 if there is none)
		const bool even = (lane & 1) == 0;
		const int u1 = lane == 31 ? 0 : u_r;
		const uint32_t evq = __ballot_sync(0xffffffffu, active && cq == 0), evq_val = __ballot_sync(0xffffffffu, rho != 0);
		const uint32_t evp = __ballot_sync(0xffffffffu, first && even && u > 0 && u1 > 0), evp_val = __ballot_sync(0xffffffffu, min(u, u1) > 2);
		// ---- VLC bits of the pair: both CxtVLC codewords, then the U-VLC prefixes and suffixes (even lanes) -------------------
		const uint32_t cwd = tuple >> 8;
		const int cl = (int) (tuple >> 4) & 7;
		const uint32_t cwd_r = __shfl_down_sync(0xffffffffu, cwd, 1);
		const int cl_r = __shfl_down_sync(0xffffffffu, cl, 1);
		uint32_t pbits = 0;
		int plen = 0;
		if (even && active) {
			uint32_t c0, c1;
			if (first && u > 2 && u1 > 2) { c0 = ht_uvlc(u - 2); c1 = ht_uvlc(u1 - 2); }
			else if (first && u > 2 && u1 > 0) { c0 = ht_uvlc(u); c1 = (uint32_t) (u1 - 1) | 1u << 8; }
			else { c0 = ht_uvlc(u); c1 = ht_uvlc(u1); }
			pbits = cwd; plen = cl;
			pbits |= cwd_r << plen; plen += cl_r;
			pbits |= (c0 & 0xFF) << plen; plen += (int) (c0 >> 8) & 0xF;
			pbits |= (c1 & 0xFF) << plen; plen += (int) (c1 >> 8) & 0xF;
			pbits |= ((c0 >> 12) & 0xFF) << plen; plen += (int) (c0 >> 20) & 0xF;
			pbits |= ((c1 >> 12) & 0xFF) << plen; plen += (int) (c1 >> 20) & 0xF;
		}
		// ---- MagSgn field widths --------------------------------------------------------------------------------------------
		int m[4], tb = 0;
		#pragma unroll
		for (int i = 0; i < 4; ++i) { m[i] = (rho >> i & 1) ? U - (int) (tuple >> i & 1) : 0; tb += m[i]; }
		// ---- bit offsets: inclusive prefix sums over the lanes (VLC in the low half, MagSgn in the high half) ---------------------
		uint32_t ps = (uint32_t) plen | (uint32_t) tb << 16;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, ps, d);
			if (lane >= d) ps += t;
		}
		const uint32_t tot = __shfl_sync(0xffffffffu, ps, 31);
		// the open bits of both writers lead the row buffers, the lanes' pieces follow
		const int vtot = (int) (tot & 0xFFFF) + vlc_used, mtot = (int) (tot >> 16) + ms_used;
		if (lane == 0) { if (vlc_used) atomicOr(vlb, vlc_acc); if (ms_used) atomicOr(msb, ms_acc); }
		ht_or_bits(vlb, (int) (ps & 0xFFFF) - plen + vlc_used, pbits, plen);
		{
			int o = (int) (ps >> 16) - tb + ms_used;
			#pragma unroll
			for (int i = 0; i < 4; ++i) { ht_or_bits(msb, o, v[i] & ((1u << m[i]) - 1u), m[i]); o += m[i]; }
		}
		__syncwarp();
		if (lane == 0) { // MEL: an adaptive run-length coder, at most three events per pair of quads, in coding order
			for (int k = 0; k < nq; k += 2) {
				if (evq >> k & 1) ht_mel_event(mel, (int) (evq_val >> k & 1));
				if (evq >> (k + 1) & 1) ht_mel_event(mel, (int) (evq_val >> (k + 1) & 1));
				if (evp >> k & 1) ht_mel_event(mel, (int) (evp_val >> k & 1));
			}
		}
		// ---- bit-stuffed byte packing, 32 candidate bytes per trip ------------------------------------------------------------
		// MagSgn: a byte is 8 bits wide, 7 after a 0xFF.  The lanes cut candidate bytes assuming no 0xFF among them; the bytes up to
		// and including the first 0xFF are right and are stored, the rest is cut again from there.
		{
			int c = 0; // cursor into the row buffer
			for (;;) {
				const int start = lane == 0 ? c : c + ms_limit + 8 * (lane - 1), width = lane == 0 ? ms_limit : 8;
				const bool full = start + width <= mtot;
				const uint32_t b = full ? ht_take_bits(msb, start, width) : 0u;
				const uint32_t have = __ballot_sync(0xffffffffu, full), ff = __ballot_sync(0xffffffffu, full && b == 0xFFu);
				if (!have) break;
				const int n = ff ? __ffs(ff) : __popc(have); // bytes of this trip that stand
				if (lane < n) { if (ms_pos + lane < ms_cap) ms_buf[ms_pos + lane] = (uint8_t) b; }
				ms_pos += n;
				c += ms_limit + 8 * (n - 1);
				ms_limit = ff ? 7 : 8;
			}
			ms_used = mtot - c;
			ms_acc = ms_used ? ht_take_bits(msb, c, ms_used) : 0u;
		}
		// VLC (written downwards): a byte that follows one > 0x8F and whose low seven bits are all ones is 7 bits wide
		{
			int c = 0;
			for (;;) {
				const int start = c + 8 * lane;
				const int avail = vtot - start;
				const uint32_t b8 = avail >= 7 ? ht_take_bits(vlb, start, min(avail, 8)) : 0u;
				uint32_t pv = __shfl_up_sync(0xffffffffu, b8, 1);
				if (lane == 0) pv = vlc_prev;
				const bool narrow = avail >= 7 && pv > 0x8Fu && (b8 & 0x7Fu) == 0x7Fu; // this byte ends after seven bits
				const bool full = avail >= 8 || narrow;
				// lanes whose byte is incomplete stop the trip, and so does the first narrow byte (the ones after it are cut wrongly)
				const uint32_t stop = __ballot_sync(0xffffffffu, !full), nar = __ballot_sync(0xffffffffu, narrow);
				int n = stop ? __ffs(stop) - 1 : 32;
				const int fn = nar ? __ffs(nar) - 1 : 32;
				const bool cut = fn < n;
				if (cut) n = fn + 1;
				if (n == 0) break;
				const uint32_t b = (cut && lane == fn) ? 0x7Fu : b8;
				if (lane < n) { if (vlc_pos + lane < HT_VLC_CAP) vlc_end[-(vlc_pos + lane)] = (uint8_t) b; }
				vlc_prev = __shfl_sync(0xffffffffu, b, n - 1);
				vlc_pos += n;
				c += 8 * n - (cut ? 1 : 0);
			}
			vlc_used = vtot - c;
			vlc_acc = vlc_used ? ht_take_bits(vlb, c, vlc_used) : 0u;
		}
		__syncwarp();
		for (int i = lane; i < ((mtot + 31) >> 5) + 1; i += 32) msb[i] = 0;
		for (int i = lane; i < ((vtot + 31) >> 5) + 1; i += 32) vlb[i] = 0;
		__syncwarp();
		pe_bl = e[1]; pe_br = e[3]; ps_bl = rho >> 1 & 1; ps_br = rho >> 3 & 1;
	}
	// ---- termination (every lane computes the same thing; lane 0 stores), then all lanes move MEL and VLC up behind MagSgn ------
	int mel_pos = 0, overflow = 0;
	{
		uint32_t mel_acc = 0;
		int mel_left = 0;
		if (lane == 0) {
			if (mel.run > 0) ht_mel_bit(mel, 1);
			mel_acc = mel.acc; mel_left = mel.left; mel_pos = mel.pos;
		}
		mel_acc = __shfl_sync(0xffffffffu, mel_acc, 0); mel_left = __shfl_sync(0xffffffffu, mel_left, 0); mel_pos = __shfl_sync(0xffffffffu, mel_pos, 0);
		const uint32_t mel_tmp = (mel_acc << mel_left) & 0xFFu;
		const uint32_t mel_mask = (0xFFu << mel_left) & 0xFFu, vlc_mask = 0xFFu >> (8 - vlc_used);
		if ((mel_mask | vlc_mask) != 0) { // the open MEL and VLC bytes share one byte when their used bits do not collide
			const uint32_t fuse = mel_tmp | vlc_acc;
			if ((((fuse ^ mel_tmp) & mel_mask) | ((fuse ^ vlc_acc) & vlc_mask)) == 0 && fuse != 0xFFu && vlc_pos > 1) {
				if (lane == 0 && mel_pos < HT_MEL_CAP) mel.buf[mel_pos] = (uint8_t) fuse;
				mel_pos++;
			} else {
				if (lane == 0 && mel_pos < HT_MEL_CAP) mel.buf[mel_pos] = (uint8_t) mel_tmp;
				mel_pos++;
				if (lane == 0 && vlc_pos < HT_VLC_CAP) vlc_end[-vlc_pos] = (uint8_t) vlc_acc;
				vlc_pos++;
			}
		}
		if (ms_used) { // pad the open MagSgn byte with ones; a padded 0xFF is dropped
			ms_acc |= ((1u << (ms_limit - ms_used)) - 1u) << ms_used;
			if (ms_acc != 0xFFu) {
				if (lane == 0 && ms_pos < ms_cap) ms_buf[ms_pos] = (uint8_t) ms_acc;
				ms_pos++;
			}
		} else if (ms_limit == 7) ms_pos--;
		overflow = ms_pos > ms_cap || mel_pos > HT_MEL_CAP || vlc_pos > HT_VLC_CAP;
	}
	const int total = ms_pos + mel_pos + vlc_pos;
	__syncwarp();
	if (!overflow) {
		// both sources lie above their destinations: chunks of 32 bytes, every lane loads before any lane stores
		const uint8_t *msrc = out + (cap - HT_MEL_CAP - HT_VLC_CAP);
		for (int i0 = 0; i0 < mel_pos; i0 += 32) {
			const int i = i0 + lane;
			const uint8_t b = i < mel_pos ? msrc[i] : 0;
			__syncwarp();
			if (i < mel_pos) out[ms_pos + i] = b;
			__syncwarp();
		}
		const uint8_t *vsrc = out + cap - vlc_pos;
		for (int i0 = 0; i0 < vlc_pos; i0 += 32) {
			const int i = i0 + lane;
			const uint8_t b = i < vlc_pos ? vsrc[i] : 0;
			__syncwarp();
			if (i < vlc_pos) out[ms_pos + mel_pos + i] = b;
			__syncwarp();
		}
		if (lane == 0) { // the last twelve bits of the block locate the MEL + VLC suffix
			const int scup = mel_pos + vlc_pos;
			out[total - 1] = (uint8_t) (scup >> 4);
			out[total - 2] = (uint8_t) ((out[total - 2] & 0xF0) | (scup & 0xF));
		}
	}
	if (lane == 0) {
		EncResult res = {1u, overflow ? 0xFFFFFFFFu : 1u, overflow ? 0u : (uint32_t) total, (uint32_t) (nq * ((h + 1) >> 1)), 0ull};
		if (B.max_passes) { rates[B.pass_offset] = res.data_len; dists[B.pass_offset] = 0.0; }
		results[bid] = res; // T1HT::encode: always one pass, numbps = 1 (T1HT.cpp:125-128)
	}
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

// ---- decoder, one WARP per code block ---------------------------------------------------------------------------------
// The MEL and VLC streams carry the significance patterns, the exponent-bound flags and the u values of every quad, and parsing
// them needs nothing from the MagSgn stream: lane 0 parses them for the whole block first (a serial chain: the context of a
// quad is its left neighbour's pattern and the row above) into one word per quad in shared memory.  The MagSgn pass is then a
// row of quads at a time with one quad per lane: kappa from the line state in registers (shuffles), field widths, a prefix sum
// for the bit offsets, and the fields are cut from a bit buffer into which the lanes have UNSTUFFED the next bytes of the stream
// (a byte's width, 7 after a 0xFF and 8 otherwise, depends only on the byte before it, so the unstuffed position of a byte is a
// prefix sum too).  De-quantisation is fused into the store.
constexpr int HTD_BUF_WORDS = HTW_MS_WORDS + 12; // a row's fields + the surplus of the last 32-byte chunk

__device__ __forceinline__ uint32_t ht_peek32(const uint32_t *buf, int pos) { // the 32 bits at bit position pos
	const int wi = pos >> 5, sh = pos & 31;
	uint32_t v = buf[wi] >> sh;
	if (sh) v |= buf[wi + 1] << (32 - sh);
	return v;
}

__global__ void __launch_bounds__(HTW_WARPS * 32) t1_ht_decode_kernel(const DecBlock *__restrict__ blocks, const DecInput *__restrict__ inputs,
		uint32_t nblocks, const uint8_t *__restrict__ data) {
	__shared__ uint32_t qinfo[HTW_WARPS][32 * 32]; // per quad: rho | e_k << 4 | e_1 << 8 | u << 12
	__shared__ uint32_t ms_bits[HTW_WARPS][HTD_BUF_WORDS + 2];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bid = blockIdx.x * HTW_WARPS + wid;
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
			for (int x = lane; x < w; x += 32) B.dst[(size_t) y * B.stride + x] = 0;
		return;
	}
	const int missing = (int) B.band_numbps - (int) I.numbps; // k_msbs (Tier1.cpp:166)
	const int p = 30 - missing;
	const int dshift = 31 - (missing + 1); // T1HT.cpp:213
	const int nq = (w + 1) >> 1, nrows = (h + 1) >> 1;
	uint32_t *Q = qinfo[wid], *msb = ms_bits[wid];
	// ---- phase 1: lane 0 parses MEL + VLC for every quad of the block ------------------------------------------------------
	if (lane == 0) {
		HtMelR mel = {D + lcup - scup, scup - 1, 0, 0, 0, 0, 0, 0, 0u};
		HtRevR vlc = {D, lcup - 3, 0, 0, 0ull};
		{
			const uint32_t d = D[lcup - 2]; // its upper nibble opens the VLC stream
			vlc.acc = d >> 4;
			vlc.bits = 4 - ((vlc.acc & 7) == 7);
			vlc.unstuff = (d | 0xF) > 0x8F;
		}
		for (int r = 0; r < nrows; ++r) {
			const bool first = r == 0;
			const uint16_t *tbl = first ? c_ht_dec0 : c_ht_dec1;
			const uint32_t *above = Q + (r - 1) * 32;
			int prev_rho = 0;
			for (int qx = 0; qx < nq; qx += 2) {
				uint32_t info[2] = {0, 0};
				#pragma unroll
				for (int k = 0; k < 2; ++k) {
					const int q = qx + k;
					if (q >= nq) break;
					int cq;
					if (first) cq = (prev_rho >> 1) | (prev_rho & 1);
					else { // bottom samples of the quads above: bit 1 = left column, bit 3 = right column
						const uint32_t a_l = q ? above[q - 1] : 0u, a = above[q], a_r = q + 1 < nq ? above[q + 1] : 0u;
						cq = (int) (((a_l >> 3) | (a >> 1)) & 1) | (((prev_rho >> 2) | (prev_rho >> 3)) & 1) << 1 | (int) (((a >> 3) | (a_r >> 1)) & 1) << 2;
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
				// table word: e_k << 12 | e_1 << 8 | rho << 4 | u_off << 3 | length
				Q[r * 32 + qx] = ((info[0] >> 4) & 15u) | ((info[0] >> 12) & 15u) << 4 | ((info[0] >> 8) & 15u) << 8 | (uint32_t) u0 << 12;
				if (qx + 1 < nq) Q[r * 32 + qx + 1] = ((info[1] >> 4) & 15u) | ((info[1] >> 12) & 15u) << 4 | ((info[1] >> 8) & 15u) << 8 | (uint32_t) u1 << 12;
			}
		}
	}
	for (int i = lane; i < HTD_BUF_WORDS + 2; i += 32) msb[i] = 0;
	__syncwarp();
	// ---- phase 2: MagSgn, a row of quads at a time, one quad per lane ---------------------------------------------------------
	const int ms_size = lcup - scup;
	int bpos = 0, boff = 0, carry = 0; // next byte of the stream, bits of it already used, whether the byte before it was 0xFF
	int pe_bl = 0, pe_br = 0;          // exponents of the bottom samples of the quad above
	const bool active = lane < nq;
	for (int r = 0; r < nrows; ++r) {
		const int y = 2 * r;
		const uint32_t qi = active ? Q[r * 32 + lane] : 0u;
		const int rho = (int) qi & 15, ek = (int) (qi >> 4) & 15, e1 = (int) (qi >> 8) & 15, u = (int) (qi >> 12);
		int kappa = 1;
		{
			int e_nw = __shfl_up_sync(0xffffffffu, pe_br, 1);
			const int e_nf = __shfl_down_sync(0xffffffffu, pe_bl, 1);
			if (lane == 0) e_nw = 0;
			if (r > 0 && (rho & (rho - 1))) kappa = max(max(max(e_nw, pe_bl), max(pe_br, lane == 31 ? 0 : e_nf)) - 1, 1);
		}
		const int Uq = u + kappa;
		int m[4], tb = 0;
		#pragma unroll
		for (int i = 0; i < 4; ++i) { m[i] = (rho >> i & 1) ? min(Uq - (ek >> i & 1), 31) : 0; tb += m[i]; } // (31: a corrupt stream cannot widen a field past the bit buffer)
		int incl = tb;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += t;
		}
		const int need = __shfl_sync(0xffffffffu, incl, 31);
		// unstuff the bytes that hold the next `need` bits into the bit buffer, 32 bytes per trip
		int got = 0, nb = bpos, nboff = boff, ncarry = carry, base = bpos, cin = carry;
		bool firstchunk = true, located = need == 0;
		while (got < need || !located) {
			const int idx = base + lane;
			const uint32_t d = idx < ms_size ? D[idx] : 0xFFu; // bytes past the end read as 0xFF
			const uint32_t dl = __shfl_up_sync(0xffffffffu, d, 1);
			const int prev_ff = lane == 0 ? cin : (dl == 0xFFu);
			const int width = 8 - prev_ff;
			int wincl = width;
			#pragma unroll
			for (int s = 1; s < 32; s <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, wincl, s);
				if (lane >= s) wincl += t;
			}
			const int skip = firstchunk ? boff : 0;          // bits of the first byte that earlier rows consumed
			const int start = got + wincl - width - skip;     // position of this byte's bit 0 in the row buffer (may be negative for the first byte)
			const int end = start + width;
			uint32_t val = d & ((1u << width) - 1u);
			int pos = start, wd = width;
			if (start < 0) { val >>= -start; wd += start; pos = 0; }
			if (wd > 0 && pos < need + 64 && pos + wd <= 32 * HTD_BUF_WORDS) ht_or_bits(msb, pos, val, wd);
			// where the stream stands after this row: the byte that holds bit `need`
			const uint32_t hit = __ballot_sync(0xffffffffu, !located && start <= need && need < end);
			if (hit) {
				const int l = __ffs(hit) - 1;
				nb = base + l;
				nboff = need - __shfl_sync(0xffffffffu, start, l);
				ncarry = __shfl_sync(0xffffffffu, prev_ff, l);
				located = true;
			}
			got += __shfl_sync(0xffffffffu, wincl, 31) - skip;
			cin = __shfl_sync(0xffffffffu, (int) (d == 0xFFu), 31);
			base += 32;
			firstchunk = false;
		}
		__syncwarp();
		// ---- the four samples of this lane's quad ------------------------------------------------------------------------------
		int o = incl - tb;
		int e_b[2] = {0, 0};
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int xx = 2 * lane + (i >> 1), yy = y + (i & 1);
			uint32_t val = 0;
			int e = 0;
			if (rho >> i & 1) {
				const uint32_t b = ht_peek32(msb, o); // bit 0 is the sign (read even when the field is empty, as the reference does)
				o += m[i];
				uint32_t vn = b & ((1u << m[i]) - 1u);
				vn |= (uint32_t) (e1 >> i & 1) << m[i]; // the implicit top bit of a sample that reaches the exponent bound
				vn |= 1;                                 // reconstruct at the centre of the bin (bit 0 was the sign)
				val = (b << 31) | ((vn + 2) << (p - 1));
				e = ht_bits(vn);
			}
			if (active && xx < w && yy < h) { // T1HT::postDecode, T1HT.cpp:211-236
				const int32_t mag = (int32_t) (val & 0x7FFFFFFFu);
				int32_t ov;
				if (B.reversible) { const int32_t q = mag >> dshift; ov = (val >> 31) ? -q : q; }
				else { const float f = __fmul_rn((float) mag, B.stepsize); ov = __float_as_int((val >> 31) ? -f : f); }
				B.dst[(size_t) yy * B.stride + xx] = ov;
			}
			if (i & 1) e_b[i >> 1] = e;
		}
		pe_bl = e_b[0]; pe_br = e_b[1];
		__syncwarp();
		for (int i = lane; i < ((max(got, need) + 31) >> 5) + 2 && i < HTD_BUF_WORDS + 2; i += 32) msb[i] = 0;
		__syncwarp();
		bpos = nb; boff = nboff; carry = ncarry;
	}
}

// ---- one THREAD per code block: the shape for LARGE batches ---------------------------------------------------------------
// With tens of thousands of blocks in flight the serial parts of the warp kernels (MEL, the VLC parse of the decoder) leave 31 lanes
// idle for most of a block's time, while 32 independent blocks per warp keep every lane busy: measured on configs[2] planes
// (49 728 blocks) the thread-per-block decoder takes 4.4 ms against 7.2 ms, on configs[1] (6 804 blocks) 2.5 ms against 1.1 ms.
// Line state in shared memory, two byte rows per thread.
__global__ void __launch_bounds__(HT_THREADS) t1_ht_encode_thread_kernel(const EncBlock *__restrict__ blocks, uint32_t nblocks,
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


__global__ void __launch_bounds__(HT_THREADS) t1_ht_decode_thread_kernel(const DecBlock *__restrict__ blocks, const DecInput *__restrict__ inputs,
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
						const int m = min(Uq - (ek >> i & 1), 31);
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

// which shape a launch takes: thread per block from this many blocks per SM on (measurement knob: the environment variable
// forces 1 = thread per block, 0 = warp per block)
#ifndef HT_DEC_THREAD_BLOCKS_PER_SM
#define HT_DEC_THREAD_BLOCKS_PER_SM 128
#endif
#ifndef HT_ENC_THREAD_BLOCKS_PER_SM
#define HT_ENC_THREAD_BLOCKS_PER_SM 100000
#endif
static bool ht_thread_per_block(uint32_t nblocks, int sms, const char *env, uint32_t per_sm) {
	const char *e = getenv(env);
	if (e && *e) return atoi(e) != 0;
	return nblocks >= (uint64_t) per_sm * (uint32_t) sms;
}

uint32_t t1_ht_scratch_extra() { return HT_MEL_CAP + HT_VLC_CAP + 64; }

void launch_t1_ht_encode(const EncBlock *blocks, uint32_t nblocks, uint8_t *scratch, EncResult *results, uint32_t *rates, double *dists,
		cudaStream_t s) {
	if (!nblocks) return;
	ensure_ht_tables();
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (ht_thread_per_block(nblocks, sms, "GB200_HT_ENC_THREAD", HT_ENC_THREAD_BLOCKS_PER_SM))
		t1_ht_encode_thread_kernel<<<(nblocks + HT_THREADS - 1) / HT_THREADS, HT_THREADS, 0, s>>>(blocks, nblocks, scratch, results, rates, dists);
	else
		t1_ht_encode_kernel<<<(nblocks + HTW_WARPS - 1) / HTW_WARPS, HTW_WARPS * 32, 0, s>>>(blocks, nblocks, scratch, results, rates, dists);
}

void launch_t1_ht_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data, cudaStream_t s) {
	if (!nblocks) return;
	ensure_ht_tables();
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (ht_thread_per_block(nblocks, sms, "GB200_HT_DEC_THREAD", HT_DEC_THREAD_BLOCKS_PER_SM))
		t1_ht_decode_thread_kernel<<<(nblocks + HT_THREADS - 1) / HT_THREADS, HT_THREADS, 0, s>>>(blocks, inputs, nblocks, data);
	else
		t1_ht_decode_kernel<<<(nblocks + HTW_WARPS - 1) / HTW_WARPS, HTW_WARPS * 32, 0, s>>>(blocks, inputs, nblocks, data);
}

} // namespace gb
