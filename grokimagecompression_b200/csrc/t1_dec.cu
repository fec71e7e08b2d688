// K5: EBCOT Tier-1 decoder with fused de-quantisation and scatter, one warp per code block.
//
//   T1Part1::decode / post_decode   T1Part1.cpp:135-329   segment concat, /2 or x stepsize, scatter
//   t1_decode_cblk                  t1.cpp:1038-1130      plane loop, pass order
//   sig / ref / cln pass            t1.cpp:381-441, 588-637, 784-870
//   MQ decoder                      mqc_dec.cpp:161-214, mqc_dec_inl.h:60-189
//
// Decoding is serial by nature (each decision selects the next context).  The warp therefore runs
// the scan as one uniform instruction stream; what the 32 lanes add is
//  * the block state as 64-bit row masks, so "which columns of this stripe can be coded at all"
//    is a handful of shifts/ORs and the scan jumps from candidate column to candidate column
//    (find-first-set) instead of visiting 64 x 4 positions,
//  * the MQ context table spread over the lanes (one shuffle per decision), the compressed bytes
//    held as a 128-byte register window across the lanes (no memory access on the decision path),
//  * magnitudes kept as one row mask per bit-plane in shared memory and turned into samples,
//    de-quantised and written coalesced by all lanes at the end (no read-modify-write of HBM).
// The reference's artificial FF FF end marker (mqc_dec.cpp:161-177) is emulated by reading 0xFF
// past the end of the segment.
#include "common.cuh"
#include "t1_tables.cuh"

namespace gb {

constexpr int DEC_WARPS = 4;

struct DecWarp {
	uint64_t sig[66], neg[66], vis[66], refd[66]; // index = row + 1
	uint64_t lastcoded[64];                       // samples coded in the plane of the last pass
};

__device__ __forceinline__ uint64_t dhor(uint64_t m) { return (m << 1) | (m >> 1); }
__device__ __forceinline__ uint64_t dfull(uint64_t m) { return m | (m << 1) | (m >> 1); }
__device__ __forceinline__ uint32_t dwin3(uint64_t m, int x) {
	return (uint32_t) (x == 0 ? (m << 1) : (m >> (x - 1))) & 7u;
}

struct MqD {
	uint32_t a, c;
	int ct;
	uint32_t pos, len;
	const uint8_t *buf;
	uint32_t wbase;  // first byte index of the register window
	uint32_t word;   // this lane's 4 bytes of the window
	uint32_t cst;    // context `lane`
};

__device__ __forceinline__ void mqd_fill(MqD &q, uint32_t base, int lane) {
	q.wbase = base;
	uint32_t w = 0;
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		uint32_t i = base + 4 * lane + j;
		uint32_t b = i < q.len ? q.buf[i] : 0xFFu;
		w |= b << (8 * j);
	}
	q.word = w;
}

__device__ __forceinline__ uint32_t mqd_byte(MqD &q, uint32_t i, int lane) {
	if (i >= q.len) return 0xFFu;
	if (i - q.wbase >= 128u) mqd_fill(q, i, lane);
	uint32_t o = i - q.wbase;
	uint32_t w = __shfl_sync(0xffffffffu, q.word, o >> 2);
	return (w >> (8 * (o & 3))) & 0xFFu;
}

__device__ __forceinline__ void mqd_bytein(MqD &q, int lane) {
	uint32_t cur = mqd_byte(q, q.pos, lane);
	uint32_t next = mqd_byte(q, q.pos + 1, lane);
	if (cur == 0xFF) {
		if (next > 0x8F) { q.c += 0xFF00u; q.ct = 8; }
		else { q.pos++; q.c += next << 9; q.ct = 7; }
	} else { q.pos++; q.c += next << 8; q.ct = 8; }
}

__device__ __forceinline__ uint32_t mqd_decode(MqD &q, uint32_t cx, int lane) {
	uint32_t st = __shfl_sync(0xffffffffu, q.cst, cx);
	uint32_t row = c_mq[st >> 1];
	uint32_t qe = row & 0xFFFFu, mps = st & 1, d;
	q.a -= qe;
	bool lps;
	if ((q.c >> 16) < qe) {
		lps = q.a >= qe; // conditional exchange
		q.a = qe;
	} else {
		q.c -= qe << 16;
		if (q.a & 0x8000u) return mps;
		lps = q.a < qe;
	}
	if (lps) { d = mps ^ 1; st = (((row >> 22) & 63u) << 1) | (mps ^ (row >> 28)); }
	else { d = mps; st = (((row >> 16) & 63u) << 1) | mps; }
	if (lane == (int) cx) q.cst = st;
	int sh = __clz(q.a) - 16;
	while (sh > 0) {
		if (q.ct == 0) mqd_bytein(q, lane);
		int n = sh < q.ct ? sh : q.ct;
		q.a <<= n; q.c <<= n; q.ct -= n; sh -= n;
	}
	return d;
}

extern __shared__ uint64_t dec_dyn_smem[];

__global__ void __launch_bounds__(DEC_WARPS * 32) t1_decode_kernel(const DecBlock *__restrict__ blocks,
		const DecInput *__restrict__ inputs, uint32_t nblocks, const uint8_t *__restrict__ data, uint32_t max_planes) {
	__shared__ DecWarp warps[DEC_WARPS];
	__shared__ uint8_t Lzc[4][256];
	__shared__ uint8_t Lsc[256];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) Lzc[i >> 8][i & 255] = c_zc[i >> 8][i & 255];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) Lsc[i] = c_sc[i];
	__syncthreads();

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bid = blockIdx.x * DEC_WARPS + wid;
	if (bid >= nblocks) return;
	DecWarp &W = warps[wid];
	uint64_t *planes = dec_dyn_smem + (size_t) wid * max_planes * 64; // [plane-1][row]
	const DecBlock B = blocks[bid];
	const DecInput I = inputs[bid];
	const int w = B.w, h = B.h;
	const uint64_t wmask = w >= 64 ? ~0ull : ((1ull << w) - 1);
	const uint8_t *zc = Lzc[B.orient];
	int numbps = (int) I.numbps;

	bool empty = I.numpasses == 0 || I.data_len == 0 || numbps == 0 || numbps > (int) max_planes; // T1Part1.cpp:139
	if (!empty) {
		for (int i = lane; i < 66; i += 32) { W.sig[i] = 0; W.neg[i] = 0; W.vis[i] = 0; W.refd[i] = 0; }
		for (int i = lane; i < 64; i += 32) W.lastcoded[i] = 0;
		for (int i = lane; i < numbps * 64; i += 32) planes[i] = 0;
		__syncwarp();

		MqD q;
		q.buf = data + I.data_offset;
		q.len = I.data_len;
		q.pos = 0;
		q.cst = lane == CTX_ZC0 ? (4 << 1) : lane == CTX_AGG ? (3 << 1) : lane == CTX_UNI ? (46 << 1) : 0;
		mqd_fill(q, 0, lane);
		q.c = mqd_byte(q, 0, lane) << 16; // INITDEC, mqc_dec.cpp:179-201
		mqd_bytein(q, lane);
		q.c <<= 7;
		q.ct -= 7;
		q.a = 0x8000;

		int bp1 = numbps, type = 2, lastplane = numbps;
		for (uint32_t pass = 0; pass < I.numpasses && bp1 >= 1; ++pass) {
			uint64_t *P = planes + (size_t) (bp1 - 1) * 64;
			lastplane = bp1;
			if (type == 0) { for (int i = lane; i < 64; i += 32) W.lastcoded[i] = 0; __syncwarp(); }
			for (int y0 = 0; y0 < h; y0 += 4) {
				const int nk = min(4, h - y0);
				uint64_t S[6], G[6], M[4], N[4], H0[4];
				#pragma unroll
				for (int j = 0; j < 6; ++j) { S[j] = W.sig[y0 + j]; G[j] = W.neg[y0 + j]; }
				uint64_t cols = 0;
				#pragma unroll
				for (int k = 0; k < 4; ++k) {
					H0[k] = dhor(S[k + 1]) | dfull(S[k]) | dfull(S[k + 2]);
					N[k] = 0;
					uint64_t valid = k < nk ? wmask : 0;
					if (type == 0) M[k] = ~S[k + 1] & H0[k] & valid;          // initial candidates
					else if (type == 1) M[k] = S[k + 1] & ~W.vis[y0 + 1 + k] & valid;
					else M[k] = ~S[k + 1] & ~W.vis[y0 + 1 + k] & valid;
					cols |= M[k];
				}
				uint64_t V[4] = {0, 0, 0, 0}; // visited in this pass (sig pass)
				uint64_t R[4] = {0, 0, 0, 0}; // magnitude bits decoded in this pass
				const uint64_t nbr_all = dfull(S[0] | S[1] | S[2] | S[3] | S[4] | S[5]);
				while (cols) {
					const int x = __ffsll((long long) cols) - 1;
					const uint64_t xb = 1ull << x;
					cols &= ~xb;
					if (type == 1) {
						#pragma unroll
						for (int k = 0; k < 4; ++k) if (M[k] & xb) {
							uint32_t ctx = (W.refd[y0 + 1 + k] & xb) ? CTX_MR0 + 2 : (H0[k] & xb) ? CTX_MR0 + 1 : CTX_MR0;
							if (mqd_decode(q, ctx, lane)) R[k] |= xb;
						}
						continue;
					}
					int k0 = 0;
					bool implied = false;
					if (type == 2 && nk == 4 && (M[0] & M[1] & M[2] & M[3] & xb) && !(nbr_all & xb)
							&& !(((N[0] | N[1] | N[2] | N[3]) << 1) & xb)) {
						if (!mqd_decode(q, CTX_AGG, lane)) continue;
						uint32_t r = mqd_decode(q, CTX_UNI, lane);
						r = (r << 1) | mqd_decode(q, CTX_UNI, lane);
						k0 = (int) r;
						implied = true;
					}
					#pragma unroll
					for (int k = 0; k < 4; ++k) {
						if (k < k0 || k >= nk) continue;
						const uint64_t up = k > 0 ? N[k - 1] : 0;
						if (type == 0) {
							if (S[k + 1] & xb) continue;
							uint64_t west = (up | N[k] | (k < 3 ? N[k + 1] : 0)) << 1;
							if (!((H0[k] | west | up) & xb)) continue;
						} else if (!(M[k] & xb)) continue;
						uint32_t top = k == 0 ? dwin3(S[0], x) : (dwin3(S[k], x) | (dwin3(N[k - 1], x) & 3));
						uint32_t mid = dwin3(S[k + 1], x) | (dwin3(N[k], x) & 1);
						uint32_t bot = k == 3 ? dwin3(S[5], x) : (dwin3(S[k + 2], x) | (dwin3(N[k + 1], x) & 1));
						uint32_t d = 1;
						if (!(implied && k == k0)) {
							uint32_t idx = top | (mid & 1) << 3 | (mid >> 2) << 4 | bot << 5;
							d = mqd_decode(q, zc[idx], lane);
						}
						if (type == 0) V[k] |= xb;
						if (d) {
							uint32_t gt = dwin3(G[k], x), gm = dwin3(G[k + 1], x), gb = dwin3(G[k + 2], x);
							uint32_t sN = top >> 1 & 1, sW = mid & 1, sE = mid >> 2 & 1, sS = bot >> 1 & 1;
							uint32_t idx = sN | sW << 1 | sE << 2 | sS << 3 | ((gt >> 1) & sN) << 4 | (gm & sW) << 5
									| ((gm >> 2) & sE) << 6 | ((gb >> 1) & sS) << 7;
							uint32_t v = Lsc[idx];
							uint32_t neg = mqd_decode(q, v & 31, lane) ^ (v >> 5);
							N[k] |= xb;
							if (neg) G[k + 1] |= xb;
							if (type == 0 && x + 1 < w) cols |= xb << 1; // east column may have become a candidate
						}
					}
				}
				// commit the stripe
				if (lane == 0) {
					#pragma unroll
					for (int k = 0; k < 4; ++k) if (k < nk) {
						const int y = y0 + k;
						if (type == 1) {
							W.refd[y + 1] |= M[k];
							P[y] |= R[k];
							W.lastcoded[y] |= M[k];
						} else {
							W.sig[y + 1] = S[k + 1] | N[k];
							W.neg[y + 1] = G[k + 1];
							P[y] |= N[k];
							W.lastcoded[y] |= N[k];
							if (type == 0) W.vis[y + 1] |= V[k];
						}
					}
				}
				__syncwarp();
			}
			if (type == 2) { for (int i = lane; i < 66; i += 32) W.vis[i] = 0; __syncwarp(); }
			if (++type == 3) { type = 0; bp1--; }
		}
		__syncwarp();
		// ---- reconstruct, de-quantise, scatter (T1Part1.cpp:216-329) -----------------------------
		for (int y = 0; y < h; ++y) {
			const uint64_t srow = W.sig[y + 1], nrow = W.neg[y + 1], lrow = W.lastcoded[y];
			for (int x = lane; x < w; x += 32) {
				int32_t v = 0;
				if (srow >> x & 1) {
					uint32_t mag = 0;
					for (int p = lastplane; p <= numbps; ++p) mag |= (uint32_t) (planes[(size_t) (p - 1) * 64 + y] >> x & 1) << p;
					mag |= (lrow >> x & 1) ? (1u << lastplane) >> 1 : 1u << lastplane;
					v = (nrow >> x & 1) ? -(int32_t) mag : (int32_t) mag;
				}
				int32_t o;
				if (B.reversible) o = v / 2;
				else o = __float_as_int(__fmul_rn((float) v, B.stepsize));
				B.dst[(size_t) y * B.stride + x] = o;
			}
		}
	} else {
		// nothing decoded: the reference leaves the zero-initialised tile buffer untouched
		for (int y = 0; y < h; ++y)
			for (int x = lane; x < w; x += 32) B.dst[(size_t) y * B.stride + x] = 0;
	}
}

static bool g_dec_tables_ready = false;

void launch_t1_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data,
		uint32_t max_planes, cudaStream_t s) {
	if (!nblocks) return;
	if (!g_dec_tables_ready) { build_and_upload_t1_tables(); g_dec_tables_ready = true; }
	if (max_planes < 1) max_planes = 1;
	size_t dyn = (size_t) DEC_WARPS * max_planes * 64 * sizeof(uint64_t);
	cudaFuncSetAttribute(t1_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dyn);
	t1_decode_kernel<<<(nblocks + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, dyn, s>>>(blocks, inputs, nblocks, data, max_planes);
}

} // namespace gb
