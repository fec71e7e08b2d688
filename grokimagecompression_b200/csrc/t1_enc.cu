// K4 (+K6): EBCOT Tier-1 encoder in two kernels, quantisation fused into the load.
//
//   T1Part1::preEncode   T1Part1.cpp:58-95    quantise (x64 or fixed-point multiply), block max
//   t1_encode_cblk       t1.cpp:1182-1326     plane loop, pass order, rates, distortion
//   sig / ref / cln pass t1.cpp:287-338, 498-555, 739-782 (steps 197-231, 443-463, 639-699)
//   MQ encoder           mqc_enc.cpp:168-287
//
// Context modelling and arithmetic coding have opposite shapes, so they are separate kernels:
//
// t1_model_kernel -- one WARP per code block (not the reference's column-serial flag-word walk):
//  * Block state is five 64-bit row masks per row in shared memory (significant, negative,
//    visited, refined, current bit-plane), built with warp ballots; a 64-bit word is one row.
//  * Which samples a pass codes is computed BIT-PARALLEL for a whole stripe (4 rows x 64 columns)
//    with shifts and ORs.  The significance-propagation pass needs the samples that became
//    significant earlier in the same pass; because the encoder knows every bit in advance this is
//    a monotone fixed point over the stripe's masks, reached in a few iterations.
//  * The 32 lanes then form contexts for 32 columns at once (8-neighbour windows cut from the
//    masks, the scan-order visibility of newly significant neighbours applied with masks) and
//    append (context, decision) bytes in scan order to the block's symbol stream (warp prefix sum,
//    coalesced stores); a marker byte closes every coding pass.  Distortion estimates are summed
//    here (they do not depend on the arithmetic coder).
//
// t1_mq_kernel -- one THREAD per code block, 32 independent MQ coders per warp: the coder is a
//  strictly serial state machine, so the only parallelism is across blocks; run this way a warp
//  instruction advances 32 streams instead of one.  Context states sit in shared memory as packed
//  Table C.2 rows ([context][thread], conflict free), symbols arrive eight at a time with the next
//  load already in flight, bytes leave with a one-byte carry delay.  Pass rates (t1.cpp:1255-1324)
//  are produced here.
#include "common.cuh"
#include <cstdlib>
#include "t1_tables.cuh"

namespace gb {

#ifndef ENC_WARPS_PER_CTA
#define ENC_WARPS_PER_CTA 4
#endif
constexpr int ENC_WARPS = ENC_WARPS_PER_CTA;
#ifndef ENC_MIN_CTAS
#define ENC_MIN_CTAS 6
#endif
constexpr int QCAP = 32 * 11 + 32;

struct EncWarp {
	uint64_t sig[66], neg[66], vis[66], refd[66], bit[66]; // index = row + 1 (rows -1 and 64 stay zero)
	uint8_t queue[QCAP];
};

struct EncLuts {
	uint8_t zc[4][256];
	uint8_t sc[256];
	int16_t nmsedec[4][128];
};

__device__ __forceinline__ uint64_t hor(uint64_t m) { return (m << 1) | (m >> 1); }
__device__ __forceinline__ uint64_t full(uint64_t m) { return m | (m << 1) | (m >> 1); }
// bits (x-1, x, x+1) of a row mask as bits 0..2
__device__ __forceinline__ uint32_t win3(uint64_t m, int x) {
	return (uint32_t) (x == 0 ? (m << 1) : (m >> (x - 1))) & 7u;
}

// marker bytes close every coding pass: 0x80 | terminated | last pass << 1 | next pass bypasses the MQ coder << 2
constexpr uint32_t SYM_PASS_END = 0x80;   // not terminated
constexpr uint32_t SYM_TERM = 1, SYM_LAST = 2, SYM_NEXT_RAW = 4;
constexpr uint32_t SYM_FLUSH_END = 0x80 | SYM_TERM | SYM_LAST; // end of the block's last pass

struct SymOut {
	uint8_t *base;   // symbol stream of this block
	uint32_t pos, cap;
	uint32_t overflow;
};

// quantised magnitude with 6 fractional bits and sign (T1Part1.cpp:45-56, 75, 85)
__device__ __forceinline__ int32_t quantise(int32_t x, bool rev, int32_t inv_step) {
	return rev ? x * 64 : (int32_t) (((int64_t) x * inv_step + (1 << 17)) >> 18);
}

// warp-wide: lanes hand in `cnt` symbols each (packed 8 bits per symbol, first symbol lowest);
// they are appended to the block's symbol stream in lane order (staged in shared memory so that the
// global stores are runs of consecutive bytes).
__device__ __forceinline__ void emit(EncWarp &W, SymOut &o, uint64_t lo, uint32_t hi, int cnt, int lane) {
	int incl = cnt;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += t;
	}
	const int total = __shfl_sync(0xffffffffu, incl, 31);
	if (total == 0) return;
	const int off = incl - cnt;
	for (int j = 0; j < cnt; ++j) {
		uint32_t s = j < 8 ? (uint32_t) (lo >> (8 * j)) & 0xFF : (hi >> (8 * (j - 8))) & 0xFF;
		W.queue[off + j] = (uint8_t) s;
	}
	__syncwarp();
	if (o.pos + (uint32_t) total <= o.cap) {
		for (int i = lane; i < total; i += 32) o.base[o.pos + i] = W.queue[i];
	} else o.overflow = 1;
	o.pos += (uint32_t) total;
	__syncwarp();
}

__device__ __forceinline__ void emit_marker(SymOut &o, uint32_t m, int lane) {
	if (o.pos < o.cap) { if (lane == 0) o.base[o.pos] = (uint8_t) m; }
	else o.overflow = 1;
	o.pos++;
}

#define PUSH(sym) do { uint32_t s_ = (sym); if (cnt < 8) lo |= (uint64_t) s_ << (8 * cnt); else hi |= s_ << (8 * (cnt - 8)); cnt++; } while (0)

__global__ void __launch_bounds__(ENC_WARPS * 32, ENC_MIN_CTAS) t1_model_kernel(const EncBlock *__restrict__ blocks, uint32_t nblocks,
		int rate_control, uint8_t *__restrict__ symbols, EncResult *__restrict__ results, double *__restrict__ dists) {
	__shared__ EncWarp warps[ENC_WARPS];
	__shared__ EncLuts L;
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) L.zc[i >> 8][i & 255] = c_zc[i >> 8][i & 255];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) L.sc[i] = c_sc[i];
	for (int i = threadIdx.x; i < 512; i += blockDim.x) L.nmsedec[i >> 7][i & 127] = c_nmsedec[i >> 7][i & 127];
	__syncthreads();

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bid = blockIdx.x * ENC_WARPS + wid;
	if (bid >= nblocks) return;
	EncWarp &W = warps[wid];
	const EncBlock B = blocks[bid];
	const int w = B.w, h = B.h;
	const bool rev = B.reversible != 0;
	const uint64_t wmask = w >= 64 ? ~0ull : ((1ull << w) - 1);
	const uint8_t *zc = L.zc[B.orient];
	const uint32_t sty = B.sty;

	// ---- quantise, block maximum, sign masks ------------------------------------------------
	for (int i = lane; i < 66; i += 32) { W.sig[i] = 0; W.neg[i] = 0; W.vis[i] = 0; W.refd[i] = 0; W.bit[i] = 0; }
	__syncwarp();
	uint32_t mx = 0;
	const bool in0 = lane < w, in1 = lane + 32 < w;
	for (int yb = 0; yb < h; yb += 8) { // eight rows per trip: all sixteen loads of the warp in flight before the first ballot
		int32_t r0[8], r1[8];
		#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int32_t *row = B.src + (size_t) (yb + j) * B.stride;
			r0[j] = (in0 && yb + j < h) ? row[lane] : 0;
			r1[j] = (in1 && yb + j < h) ? row[lane + 32] : 0;
		}
		#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int32_t v0 = quantise(r0[j], rev, B.inv_step), v1 = quantise(r1[j], rev, B.inv_step);
			mx = max(mx, (uint32_t) abs(v0));
			mx = max(mx, (uint32_t) abs(v1));
			const uint32_t n0 = __ballot_sync(0xffffffffu, v0 < 0), n1 = __ballot_sync(0xffffffffu, v1 < 0);
			if (lane == j && yb + j < h) W.neg[yb + j + 1] = (uint64_t) n0 | ((uint64_t) n1 << 32);
		}
	}
	#pragma unroll
	for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	const int nbits = 32 - __clz(mx);
	const int numbps = nbits > 6 ? nbits - 6 : 0; // t1.cpp:1202-1208
	if (numbps == 0) {
		if (lane == 0) { EncResult r = {0, 0, 0, 0, 0}; results[bid] = r; }
		return;
	}

	SymOut q;
	q.base = symbols + B.sym_off;
	q.pos = 0; q.cap = B.sym_cap; q.overflow = 0;

	double *my_dists = dists + B.pass_offset;
	int npass = 0;
	double cum = 0.0;

	for (int bp = numbps - 1; bp >= 0; --bp) {
		// ---- bit-plane masks for this plane -------------------------------------------------
		for (int yb = 0; yb < h; yb += 8) {
			int32_t r0[8], r1[8];
			#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const int32_t *row = B.src + (size_t) (yb + j) * B.stride;
				r0[j] = (in0 && yb + j < h) ? row[lane] : 0;
				r1[j] = (in1 && yb + j < h) ? row[lane + 32] : 0;
			}
			#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const uint32_t m0 = (uint32_t) abs(quantise(r0[j], rev, B.inv_step)), m1 = (uint32_t) abs(quantise(r1[j], rev, B.inv_step));
				const uint32_t b0 = __ballot_sync(0xffffffffu, (m0 >> (bp + 6)) & 1), b1 = __ballot_sync(0xffffffffu, (m1 >> (bp + 6)) & 1);
				if (lane == j && yb + j < h) W.bit[yb + j + 1] = (uint64_t) b0 | ((uint64_t) b1 << 32);
			}
		}
		__syncwarp();

		for (int type = (bp == numbps - 1 ? 2 : 0); type < 3; ++type) {
			int nmsedec = 0;
			// selective arithmetic coding bypass: raw passes carry the sign itself (t1.cpp:224-229, 1229-1231)
			const bool raw = (sty & STY_LAZY) && bp < numbps - 4 && type < 2;
			for (int y0 = 0; y0 < h; y0 += 4) {
				const int nk = min(4, h - y0);
				uint64_t S[6], Bm[4], M[4], N[4], H0[4];
				#pragma unroll
				for (int j = 0; j < 6; ++j) S[j] = W.sig[y0 + j];
				if (sty & STY_VSC) S[5] = 0; // stripe-causal contexts: the row below the stripe is invisible (t1.cpp:177-182)
				#pragma unroll
				for (int k = 0; k < 4; ++k) {
					Bm[k] = W.bit[y0 + 1 + k];
					H0[k] = hor(S[k + 1]) | full(S[k]) | full(S[k + 2]);
					N[k] = 0;
				}
				uint64_t rl = 0;
				if (type == 0) {
					// membership of the significance-propagation pass: least fixed point
					bool changed = true;
					while (changed) {
						changed = false;
						#pragma unroll
						for (int k = 0; k < 4; ++k) {
							uint64_t up = k > 0 ? N[k - 1] : 0;
							uint64_t west = (up | N[k] | (k < 3 ? N[k + 1] : 0)) << 1;
							M[k] = k < nk ? (~S[k + 1] & (H0[k] | west | up) & wmask) : 0;
							uint64_t n = M[k] & Bm[k];
							if (n != N[k]) { changed = true; N[k] = n; }
						}
					}
				} else if (type == 1) {
					#pragma unroll
					for (int k = 0; k < 4; ++k) M[k] = k < nk ? (S[k + 1] & ~W.vis[y0 + 1 + k] & wmask) : 0;
				} else {
					#pragma unroll
					for (int k = 0; k < 4; ++k) {
						M[k] = k < nk ? (~S[k + 1] & ~W.vis[y0 + 1 + k] & wmask) : 0;
						N[k] = M[k] & Bm[k];
					}
					if (nk == 4)
						rl = M[0] & M[1] & M[2] & M[3] & ~full(S[0] | S[1] | S[2] | S[3] | S[4] | S[5])
								& ~((N[0] | N[1] | N[2] | N[3]) << 1);
				}
				const uint64_t any = M[0] | M[1] | M[2] | M[3];
				if (any) {
					for (int ch = 0; ch * 32 < w; ++ch) {
						if (((any >> (32 * ch)) & 0xffffffffull) == 0) continue;
						const int gx = ch * 32 + lane;
						uint64_t lo = 0; uint32_t hi = 0; int cnt = 0;
						uint32_t mem4 = 0;
						#pragma unroll
						for (int k = 0; k < 4; ++k) mem4 |= (uint32_t) ((M[k] >> gx) & 1) << k;
						if (mem4) {
							if (type == 1) {
								#pragma unroll
								for (int k = 0; k < 4; ++k) if (mem4 >> k & 1) {
									uint32_t ctx = (W.refd[y0 + 1 + k] >> gx & 1) ? CTX_MR0 + 2 : (H0[k] >> gx & 1) ? CTX_MR0 + 1 : CTX_MR0;
									PUSH(ctx << 1 | (uint32_t) (Bm[k] >> gx & 1));
									if (rate_control) {
										uint32_t mag = (uint32_t) abs(quantise(B.src[(size_t) (y0 + k) * B.stride + gx], rev, B.inv_step));
										nmsedec += L.nmsedec[bp > 0 ? 2 : 3][(mag >> bp) & 127];
									}
								}
							} else {
								uint32_t sw[6], nw[4], gw[6];
								#pragma unroll
								for (int j = 0; j < 6; ++j) { sw[j] = win3(S[j], gx); gw[j] = win3(W.neg[y0 + j], gx); }
								#pragma unroll
								for (int k = 0; k < 4; ++k) nw[k] = win3(N[k], gx);
								int k0 = 0;
								bool implied = false; // first sample after a run: its 1 is implied, go straight to the sign
								if (type == 2 && (rl >> gx & 1)) {
									uint32_t b4 = (uint32_t) (Bm[0] >> gx & 1) | (uint32_t) (Bm[1] >> gx & 1) << 1
											| (uint32_t) (Bm[2] >> gx & 1) << 2 | (uint32_t) (Bm[3] >> gx & 1) << 3;
									int r = b4 ? __ffs(b4) - 1 : 4;
									PUSH(CTX_AGG << 1 | (r != 4));
									if (r == 4) k0 = 4;
									else {
										PUSH(CTX_UNI << 1 | (r >> 1));
										PUSH(CTX_UNI << 1 | (r & 1));
										k0 = r;
										implied = true;
									}
								}
								#pragma unroll
								for (int k = 0; k < 4; ++k) {
									if (k < k0 || !(mem4 >> k & 1)) continue;
									// 8-neighbourhood as visible when the scan reaches (gx, y0+k): newly
									// significant samples count only in the west column and above in this column
									uint32_t top = k == 0 ? sw[0] : (sw[k] | (nw[k - 1] & 3));
									uint32_t mid = sw[k + 1] | (nw[k] & 1);
									uint32_t bot = k == 3 ? sw[5] : (sw[k + 2] | (nw[k + 1] & 1));
									uint32_t d = (uint32_t) (Bm[k] >> gx & 1);
									if (!(implied && k == k0)) {
										uint32_t idx = top | (mid & 1) << 3 | (mid >> 2) << 4 | bot << 5;
										PUSH((uint32_t) zc[idx] << 1 | d);
									}
									if (d) {
										uint32_t sN = top >> 1 & 1, sW = mid & 1, sE = mid >> 2 & 1, sS = bot >> 1 & 1;
										uint32_t gN = gw[k] >> 1 & 1, gW = gw[k + 1] & 1, gE = gw[k + 1] >> 2 & 1, gS = gw[k + 2] >> 1 & 1;
										uint32_t idx = sN | sW << 1 | sE << 2 | sS << 3 | (gN & sN) << 4 | (gW & sW) << 5 | (gE & sE) << 6 | (gS & sS) << 7;
										uint32_t v = L.sc[idx];
										uint32_t sgn = gw[k + 1] >> 1 & 1;
										PUSH((v & 31) << 1 | (raw ? sgn : sgn ^ (v >> 5)));
										if (rate_control) {
											uint32_t mag = (uint32_t) abs(quantise(B.src[(size_t) (y0 + k) * B.stride + gx], rev, B.inv_step));
											nmsedec += L.nmsedec[bp > 0 ? 0 : 1][(mag >> bp) & 127];
										}
									}
								}
							}
						}
						emit(W, q, lo, hi, cnt, lane);
					}
				}
				// commit the stripe
				if (lane == 0) {
					#pragma unroll
					for (int k = 0; k < 4; ++k) if (k < nk) {
						if (type != 1) W.sig[y0 + 1 + k] |= N[k];
						if (type == 0) W.vis[y0 + 1 + k] |= M[k];
						if (type == 1) W.refd[y0 + 1 + k] |= M[k];
					}
				}
				__syncwarp();
			}
			if (type == 2) {
				for (int i = lane; i < 66; i += 32) W.vis[i] = 0;
				__syncwarp();
				if (sty & STY_SEGSYM) // segmentation symbol 1010 in the UNIFORM context (mqc_enc.cpp:409-413)
					emit(W, q, lane == 0 ? (uint64_t) (CTX_UNI << 1 | 1) | (uint64_t) (CTX_UNI << 1) << 8 | (uint64_t) (CTX_UNI << 1 | 1) << 16
							| (uint64_t) (CTX_UNI << 1) << 24 : 0ull, 0u, lane == 0 ? 4 : 0, lane);
			}
			// ---- pass bookkeeping (t1.cpp:1255-1290) ----------------------------------------
			if (rate_control) {
				#pragma unroll
				for (int o = 16; o; o >>= 1) nmsedec += __shfl_xor_sync(0xffffffffu, nmsedec, o);
				double x = __dmul_rn(B.rd_weight, (double) (1 << bp));
				x = __dmul_rn(x, __ddiv_rn(__dmul_rn(x, (double) nmsedec), 8192.0));
				cum = __dadd_rn(cum, x);
			}
			{ // t1_enc_is_term_pass, t1.cpp:1131-1151
				const bool last = type == 2 && bp == 0;
				bool term = last || (sty & STY_TERMALL);
				if (sty & STY_LAZY) term = term || (bp == numbps - 4 && type == 2) || (bp < numbps - 4 && type > 0);
				const int ntype = type == 2 ? 0 : type + 1, nbp = type == 2 ? bp - 1 : bp;
				const bool next_raw = !last && (sty & STY_LAZY) && nbp < numbps - 4 && ntype < 2;
				emit_marker(q, SYM_PASS_END | (term ? SYM_TERM : 0u) | (last ? SYM_LAST : 0u) | (next_raw ? SYM_NEXT_RAW : 0u), lane);
			}
			if (lane == 0 && (uint32_t) npass < B.max_passes) my_dists[npass] = rate_control ? cum : 0.0;
			npass++;
		}
	}
	if (lane == 0) {
		EncResult res;
		res.numbps = (uint32_t) numbps;
		res.numpasses = (q.overflow || npass > (int) B.max_passes) ? 0xFFFFFFFFu : (uint32_t) npass;
		res.data_len = 0;
		res.decisions = 0;
		res.data_offset = 0;
		results[bid] = res;
	}
}

// ---- MQ coder: one thread per code block ------------------------------------------------------------
// Lean scalar code, MQ_LANES coders per warp: what bounds this kernel is the number of instructions of one
// coder's serial chain, so the step is written the way the reference's CODEMPS / CODELPS / RENORME are
// (mqc_enc.cpp:168-243), with the SWITCH column folded into a 94-entry (state, mps) table and A kept in the
// high half-word so that one CLZ gives the renormalisation shift.

#ifndef MQ_WARPS_PER_CTA
#define MQ_WARPS_PER_CTA 4
#endif
constexpr int MQ_WARPS = MQ_WARPS_PER_CTA;
#ifndef MQ_SPARSE_BLOCKS_PER_SM
#define MQ_SPARSE_BLOCKS_PER_SM 16
#endif
constexpr int MQ_CTX_WORDS = 20;
#ifndef MQ_PREFETCH
#define MQ_PREFETCH 1
#endif

struct MqT {
	uint32_t a, c;     // A kept in the high half-word (a << 16) so that the renormalisation shift is clz(a)
	int ct;
	int pos;           // index of the byte held in `last`; -1 = the pad byte in front of the stream
	uint32_t last;
};

__device__ __forceinline__ void mqt_byteout(MqT &q, uint8_t *out, uint32_t cap, uint32_t &overflow) {
	if (q.last != 0xFF && (q.c & 0x8000000u)) { // carry into the delayed byte
		q.last++;
		q.c &= 0x7FFFFFFu;
	}
	if (q.pos >= 0) {
		if ((uint32_t) q.pos < cap) out[q.pos] = (uint8_t) q.last; else overflow = 1;
	}
	q.pos++;
	if (q.last == 0xFF) { q.last = (q.c >> 20) & 0xFF; q.c &= 0xFFFFFu; q.ct = 7; }
	else { q.last = (q.c >> 19) & 0xFF; q.c &= 0x7FFFFu; q.ct = 8; }
}

template<int LANES, bool STY>
__global__ void __launch_bounds__(MQ_WARPS * 32) t1_mq_kernel(const EncBlock *__restrict__ blocks, uint32_t nblocks,
		const uint8_t *__restrict__ symbols, uint8_t *__restrict__ scratch, EncResult *__restrict__ results,
		uint32_t *__restrict__ rates) {
	// context rows: qe << 16 | next(LPS) << 9 | next(MPS) << 2 | mps; successors are byte offsets into the (state, mps) table
	__shared__ uint32_t tab[96];
	__shared__ uint32_t ctx[MQ_WARPS * LANES][MQ_CTX_WORDS];
	for (int i = threadIdx.x; i < 94; i += blockDim.x) {
		const uint32_t r = c_mq[i >> 1], mps = i & 1u, sw = (r >> 28) & 1u;
		const uint32_t nm = ((r >> 16) & 63u) * 2u + mps, nl = ((r >> 22) & 63u) * 2u + (mps ^ sw);
		tab[i] = (r << 16) | (nl << 9) | (nm << 2) | mps;
	}
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane >= LANES) return;
	int slot = warp * LANES + lane;
	const uint32_t bid = blockIdx.x * (MQ_WARPS * LANES) + (uint32_t) slot;
	if (bid >= nblocks) return;
	EncResult res = results[bid];
	if (res.numbps == 0 || res.numpasses == 0 || res.numpasses == 0xFFFFFFFFu) return;
	const EncBlock B = blocks[bid];
	asm volatile("" : "+r"(slot)); // keep the context base in a register instead of re-deriving it from the thread index per symbol
	uint32_t *C = ctx[slot];
	#pragma unroll
	for (int i = 0; i < NCTX; ++i) C[i] = tab[2 * (i == CTX_ZC0 ? 4 : i == CTX_AGG ? 3 : i == CTX_UNI ? 46 : 0)]; // mqc_dec.cpp:207-214
	uint8_t *out = scratch + B.scratch_off + 1;
	const uint32_t cap = B.scratch_cap - 1;
	uint32_t *my_rates = rates + B.pass_offset;
	uint32_t overflow = 0, nsym = 0;
	int npass = 0;
	MqT q;
	q.a = 0x80000000u; q.c = 0; q.ct = 12; q.pos = -1; q.last = 0;

	// CODEMPS / CODELPS + RENORME for one (context, decision) byte
	uint32_t tab_s = (uint32_t) __cvta_generic_to_shared(tab);
	asm volatile("" : "+r"(tab_s)); // the shared-window address of the table stays in a register
	auto tabrow = [&](uint32_t off) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tab_s + off)); return v; };
	// `row` is the context's row as loaded by the caller, `pre` / `pcr` a row loaded ahead for the NEXT symbol and where it came
	// from: when this symbol rewrites that very context, the row in flight is replaced by the new one
	auto code_row = [&](uint32_t sym, uint32_t *cr, uint32_t row, uint32_t &pre, const uint32_t *pcr) {
		const uint32_t qs = row & 0xFFFF0000u;
		const uint32_t a = q.a - qs;
		const bool ismps = !((row ^ sym) & 1u);
		if (ismps && (a & 0x80000000u)) { // CODEMPS without renormalisation: the one early exit
			q.a = a;
			q.c += qs >> 16;
			return;
		}
		// CODEMPS: A < Qe ? A = Qe : C += Qe;  CODELPS: A < Qe ? C += Qe : A = Qe  -- one select pair
		const bool takeq = ismps == (a < qs);
		q.a = takeq ? qs : a;
		q.c += takeq ? 0u : qs >> 16;
		const uint32_t nrow = tabrow(ismps ? row & 0x1FCu : (row >> 7) & 0x1FCu);
		*cr = nrow;
		if (pcr == cr) pre = nrow;
		int sh = __clz(q.a); // RENORME
		q.a <<= sh;
		while (sh >= q.ct) { // a byte is completed inside this shift
			q.c <<= q.ct;
			sh -= q.ct;
			mqt_byteout(q, out, cap, overflow);
		}
		q.c <<= sh;
		q.ct -= sh;
	};
	auto code = [&](uint32_t sym) {
		uint32_t *cr = C + (sym >> 1);
		uint32_t none = 0;
		code_row(sym, cr, *cr, none, nullptr);
	};

	const uint2 *sp = reinterpret_cast<const uint2*>(symbols + B.sym_off); // sym_off is 16-byte aligned
	uint2 nxt = sp[0];
	uint32_t word = 1, consumed = 0;
	bool done = false;
	if (!STY) {
		// one byte per symbol, a marker closes every coding pass; eight symbols per load, the next eight already in
		// flight.  Groups without a marker (all but ~1 in 200) run fully unrolled.
		while (!done) {
			const uint2 cur = nxt;
			nxt = sp[word++];
			if (((cur.x | cur.y) & 0x80808080u) == 0) {
#if MQ_PREFETCH
				// the context row of symbol j + 1 is loaded before symbol j is coded: the shared-memory latency leaves the chain
				uint32_t sym = cur.x & 0xFFu;
				uint32_t *cr = C + (sym >> 1);
				uint32_t row = *cr;
				#pragma unroll
				for (int j = 0; j < 8; ++j) {
					const uint32_t nsym = j < 7 ? ((j + 1 < 4 ? cur.x : cur.y) >> (8 * ((j + 1) & 3))) & 0xFFu : 0u;
					uint32_t *ncr = C + (nsym >> 1);
					uint32_t nrow = j < 7 ? *ncr : 0u;
					code_row(sym, cr, row, nrow, j < 7 ? ncr : nullptr);
					sym = nsym; cr = ncr; row = nrow;
				}
#else
				#pragma unroll
				for (int j = 0; j < 8; ++j) code(((j < 4 ? cur.x : cur.y) >> (8 * (j & 3))) & 0xFFu);
#endif
			} else {
				uint64_t grp = (uint64_t) cur.x | ((uint64_t) cur.y << 32);
				#pragma unroll 1
				for (uint32_t j = 0; j < 8 && !done; ++j, grp >>= 8) {
					const uint32_t sym = (uint32_t) grp & 0xFFu;
					if (!(sym & 0x80u)) { code(sym); continue; }
					// end of a coding pass (t1.cpp:1255-1290)
					uint32_t rate;
					if (sym & SYM_LAST) { // FLUSH, mqc_enc.cpp:235-243, 274-287
						const uint32_t t = q.c + (q.a >> 16);
						q.c |= 0xFFFFu;
						if (q.c >= t) q.c -= 0x8000u;
						q.c <<= q.ct; mqt_byteout(q, out, cap, overflow);
						q.c <<= q.ct; mqt_byteout(q, out, cap, overflow);
						if (q.last != 0xFF) {
							if ((uint32_t) q.pos < cap) out[q.pos] = (uint8_t) q.last; else overflow = 1;
							q.pos++;
						}
						rate = (uint32_t) q.pos;
						nsym = consumed + j - (uint32_t) npass;
						done = true;
					} else rate = (uint32_t) q.pos + (q.ct < 5 ? 6 : 5);
					if ((uint32_t) npass < B.max_passes) my_rates[npass] = rate;
					npass++;
				}
			}
			consumed += 8;
		}
	} else {
		// code-block style switches (t1.cpp:1223-1298): per-pass termination (TERMALL, LAZY), predictable termination
		// (PTERM), context reset (RESET), raw passes (LAZY).  In raw mode q.pos is the next free byte and q.c / q.ct
		// the partial byte (mqc_enc.cpp:291-377); in MQ mode q.pos is the index of the pending byte q.last.
		const uint32_t sty = B.sty;
		const bool pterm = (sty & STY_PTERM) != 0;
		constexpr int RAW_CT_INIT = 0x7FFFFFFF;
		bool raw = false;
		auto byte_at = [&](int i) -> uint32_t { return (i >= 0 && (uint32_t) i < cap) ? out[i] : 0u; };
		auto put = [&](uint32_t v) { if ((uint32_t) q.pos < cap) out[q.pos] = (uint8_t) v; else overflow = 1; q.pos++; };
		auto raw_pending = [&]() { return q.ct < 7 || (q.ct == 7 && (pterm || byte_at(q.pos - 1) != 0xFFu)); };
		while (!done) {
			const uint2 cur = nxt;
			nxt = sp[word++];
			uint64_t grp = (uint64_t) cur.x | ((uint64_t) cur.y << 32);
			#pragma unroll 1
			for (uint32_t j = 0; j < 8 && !done; ++j, grp >>= 8) {
				const uint32_t sym = (uint32_t) grp & 0xFFu;
				if (!(sym & 0x80u)) {
					if (!raw) { code(sym); continue; }
					if (q.ct == RAW_CT_INIT) q.ct = 8; // mqc_bypass_enc
					q.ct--;
					q.c += (sym & 1u) << q.ct;
					if (q.ct == 0) {
						put(q.c);
						q.ct = q.c == 0xFFu ? 7 : 8; // the byte after 0xFF keeps its MSB clear
						q.c = 0;
					}
					continue;
				}
				uint32_t rate;
				if (sym & SYM_TERM) {
					if (raw) { // mqc_bypass_flush_enc
						if (raw_pending()) {
							uint32_t bit = 0;
							while (q.ct > 0) { q.ct--; q.c += bit << q.ct; bit ^= 1u; }
							put(q.c);
						} else if (q.ct == 7 && byte_at(q.pos - 1) == 0xFFu) q.pos--;
						else if (q.ct == 8 && !pterm && byte_at(q.pos - 1) == 0x7Fu && byte_at(q.pos - 2) == 0xFFu) q.pos -= 2;
					} else if (pterm) { // mqc_erterm_enc
						int k = 11 - q.ct + 1;
						while (k > 0) {
							q.c <<= q.ct;
							q.ct = 0;
							mqt_byteout(q, out, cap, overflow);
							k -= q.ct;
						}
						if (q.last != 0xFF) mqt_byteout(q, out, cap, overflow);
					} else { // mqc_flush_enc
						const uint32_t t = q.c + (q.a >> 16);
						q.c |= 0xFFFFu;
						if (q.c >= t) q.c -= 0x8000u;
						q.c <<= q.ct; mqt_byteout(q, out, cap, overflow);
						q.c <<= q.ct; mqt_byteout(q, out, cap, overflow);
						if (q.last != 0xFF) put(q.last);
					}
					rate = (uint32_t) max(q.pos, 0);
				} else if (raw) rate = (uint32_t) q.pos + (raw_pending() ? 2u : 1u);
				else rate = (uint32_t) q.pos + (q.ct < 5 ? 6 : 5);
				if ((uint32_t) npass < B.max_passes) my_rates[npass] = rate;
				npass++;
				if (sty & STY_RESET) {
					#pragma unroll
					for (int i = 0; i < NCTX; ++i) C[i] = tab[2 * (i == CTX_ZC0 ? 4 : i == CTX_AGG ? 3 : i == CTX_UNI ? 46 : 0)];
				}
				if (sym & SYM_LAST) {
					nsym = consumed + j - (uint32_t) (npass - 1);
					done = true;
				} else if (sym & SYM_TERM) { // the next pass starts a new codeword segment
					raw = (sym & SYM_NEXT_RAW) != 0;
					if (raw) { q.c = 0; q.ct = RAW_CT_INIT; } // mqc_bypass_init_enc: q.pos is already the next free byte
					else { // mqc_restart_init_enc: the last byte of the previous segment is pending again
						q.a = 0x80000000u; q.c = 0; q.ct = 12;
						q.pos--;
						q.last = byte_at(q.pos);
						if (q.last == 0xFFu) q.ct = 13;
					}
				}
			}
			consumed += 8;
		}
	}
	// ---- rate fix-ups (t1.cpp:1300-1324): non-increasing from the end, no trailing 0xFF ------
	const int np = min(npass, (int) B.max_passes);
	uint32_t lastr = (uint32_t) q.pos;
	for (int i = np - 1; i >= 0; --i) {
		uint32_t r = my_rates[i];
		if (r > lastr) { r = lastr; my_rates[i] = r; } else lastr = r;
	}
	for (int i = 0; i < np; ++i) {
		const uint32_t r = my_rates[i];
		const uint8_t prev = r >= 1 && r - 1 < cap ? out[r - 1] : 0;
		if (prev == 0xFF) my_rates[i] = r - 1;
	}
	res.numpasses = (overflow || npass > (int) B.max_passes || npass != (int) res.numpasses) ? 0xFFFFFFFFu : (uint32_t) npass;
	res.data_len = np ? my_rates[np - 1] : 0;
	res.decisions = nsym;
	results[bid] = res;
}

// ---- compaction: exclusive prefix sum of the block lengths, then one warp copies each block ----
__global__ void __launch_bounds__(1024) t1_offsets_kernel(EncResult *results, uint32_t nblocks, uint64_t *total) {
	__shared__ uint64_t part[1024];
	const uint32_t per = (nblocks + 1023) / 1024;
	const uint32_t b0 = threadIdx.x * per, b1 = min(nblocks, b0 + per);
	uint64_t s = 0;
	for (uint32_t i = b0; i < b1; ++i) s += results[i].data_len;
	part[threadIdx.x] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint64_t run = 0;
		for (int i = 0; i < 1024; ++i) { uint64_t t = part[i]; part[i] = run; run += t; }
	}
	__syncthreads();
	uint64_t off = part[threadIdx.x];
	for (uint32_t i = b0; i < b1; ++i) { results[i].data_offset = off; off += results[i].data_len; }
	if (total && threadIdx.x == 1023) *total = off; // the last thread ends at the sum of all lengths
}

__global__ void __launch_bounds__(256) t1_gather_kernel(const EncBlock *__restrict__ blocks, const EncResult *__restrict__ results,
		uint32_t nblocks, const uint8_t *__restrict__ scratch, uint8_t *__restrict__ data) {
	const uint32_t bid = blockIdx.x * 8 + (threadIdx.x >> 5);
	if (bid >= nblocks) return;
	const int lane = threadIdx.x & 31;
	const uint8_t *src = scratch + blocks[bid].scratch_off + 1;
	uint8_t *dst = data + results[bid].data_offset;
	const uint32_t n = results[bid].data_len;
	for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

void launch_t1_encode(const EncBlock *blocks, uint32_t nblocks, int rate_control, int styles, uint8_t *symbols, uint8_t *scratch,
		EncResult *results, uint32_t *rates, double *dists, cudaStream_t s) {
	if (!nblocks) return;
	ensure_t1_tables();
	t1_model_kernel<<<(nblocks + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, s>>>(blocks, nblocks, rate_control, symbols, results, dists);
	// Coders per warp.  Few blocks per SM: the launch is bound by the latency of one coder's chain, every coder gets its own warp.
	// Otherwise the kernel is bound by warp instructions issued, and the coders of a warp share most of theirs (one uniform
	// loop, 1.6 of 2 lanes live on average): the more blocks a launch has per SM, the more coders a warp can carry without
	// starving the SM of warps.  Measured (profiles/README.md), model + MQ time: configs[1], 46 blocks per SM: 6.22 ms at 2 coders
	// per warp, 5.63 at 4, 5.86 at 6, 6.71 at 16; configs[2] planes, 336 per SM: 89 / 53 / 45 / 42 ms at 2 / 8 / 16 / 32; 30 cinema
	// frames, 1360 per SM: 58 / 38 / 34 / 31 ms.
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint32_t per_sm = (nblocks + (uint32_t) sms - 1) / (uint32_t) sms;
	int lanes = styles ? 1 : per_sm <= MQ_SPARSE_BLOCKS_PER_SM ? 1 : per_sm <= 64 ? 4 : per_sm <= 160 ? 8 : per_sm <= 320 ? 16 : 32;
	if (const char *e = getenv("GB200_T1_MQ_LANES")) { const int l = atoi(e); if (!styles && (l == 1 || l == 2 || l == 4 || l == 8 || l == 16 || l == 32)) lanes = l; } // measurement knob
	const uint32_t per_cta = (uint32_t) (MQ_WARPS * lanes), grid = (nblocks + per_cta - 1) / per_cta;
	if (styles) t1_mq_kernel<1, true><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else if (lanes == 1) t1_mq_kernel<1, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else if (lanes == 2) t1_mq_kernel<2, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else if (lanes == 4) t1_mq_kernel<4, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else if (lanes == 8) t1_mq_kernel<8, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else if (lanes == 16) t1_mq_kernel<16, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
	else t1_mq_kernel<32, false><<<grid, MQ_WARPS * 32, 0, s>>>(blocks, nblocks, symbols, scratch, results, rates);
}

// bytes of symbol stream to reserve for a w x h block with at most `planes` coded bit-planes: every sample
// yields at most one decision per plane plus one sign, run-length mode adds at most two per stripe column,
// one marker per pass; rounded up so that streams stay 16-byte aligned and 16 bytes can be read past the end
uint32_t t1_symbol_capacity(uint32_t w, uint32_t h, uint32_t planes) {
	uint64_t n = (uint64_t) planes * (w * h + ((h + 3) / 4) * w * 2) + (uint64_t) w * h + 3 * planes + 4 * planes /* SEGSYM */ + 8;
	return (uint32_t) ((n + 16 + 15) / 16 * 16);
}

void launch_t1_gather(const EncBlock *blocks, EncResult *results, uint32_t nblocks, const uint8_t *scratch, uint8_t *data,
		uint64_t *total, cudaStream_t s) {
	if (!nblocks) return;
	t1_offsets_kernel<<<1, 1024, 0, s>>>(results, nblocks, total);
	t1_gather_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(blocks, results, nblocks, scratch, data);
}

} // namespace gb
