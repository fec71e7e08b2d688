// K5: EBCOT Tier-1 decoder with fused de-quantisation and scatter, one warp per code block.
//
//   T1Part1::decode / post_decode   T1Part1.cpp:135-329   segment concat, /2 or x stepsize, scatter
//   t1_decode_cblk                  t1.cpp:1038-1130      plane loop, pass order
//   sig / ref / cln pass            t1.cpp:381-441, 588-637, 784-870
//   MQ decoder                      mqc_dec.cpp:161-214, mqc_dec_inl.h:60-189
//
// Decoding is serial by nature (each decision selects the next context), so the warp runs the scan
// as ONE uniform instruction stream and the design goal is the shortest possible dependent chain
// per decision with few live registers (occupancy hides the rest):
//  * State of a stripe column is one 32-bit word: significance of the 3x6 neighbourhood, signs of
//    the own column, visited and refined bits.  The 64 words of the current stripe live in the
//    lanes' registers (lane l owns columns l and l+32); the scan fetches a column's word with one
//    shuffle, works on it with 32-bit logic, and the lanes owning the west/east columns update
//    their own copies when a sample turns significant.  Other stripes wait in shared memory.
//  * Warp ballots turn the per-lane words into the list of columns that can be coded in this pass
//    at all; the scan jumps between them with find-first-set instead of visiting 64 x 4 positions.
//  * The MQ probability state of context i lives in lane i as the packed table row, so a decision
//    costs one shuffle and no table access on its critical path; compressed bytes are held as a
//    128-byte register window across the lanes.
//  * Decoded magnitude bits are collected as per-lane nibbles, turned into row masks with ballots
//    at the end of a stripe, kept per bit-plane, and only at the very end expanded to samples,
//    de-quantised and written coalesced (no read-modify-write of the coefficient plane).
// The reference's artificial FF FF end marker (mqc_dec.cpp:161-177) is emulated by reading 0xFF
// past the end of the segment.
#include "common.cuh"
#include "t1_tables.cuh"

namespace gb {

constexpr int DEC_WARPS = 4;
#ifndef DEC_MIN_CTAS
#define DEC_MIN_CTAS 8
#endif

// stripe-column word: bit 3r+j = significance of row r-1 (r = 0..5), column j-1 (j = 0 west, 1 own, 2 east);
// bit 18+r = sign of own column row r-1; bit 24+k = visited (k = 0..3); bit 28+k = refined before
__device__ __forceinline__ constexpr uint32_t fsig(int r, int j) { return 1u << (3 * r + j); }
constexpr uint32_t F_PI_ALL = 0xFu << 24;

struct DecWarp {
	uint32_t F[18][64];   // [stripe + 1][column]; stripes -1 and 16 are never read as a current stripe
	uint64_t pcur[64];    // magnitude bits of the current bit-plane, one row mask per row
	uint64_t lastc[64];   // samples coded in the final (possibly partial) plane
};

struct MqD {
	uint32_t a, c;   // A is kept in the high half-word (A << 16), like the Chigh half of C it is compared with
	int ct;
	uint32_t pos, len;
	const uint8_t *buf;
	uint32_t wbase;  // first byte index of the register window
	uint32_t word;   // this lane's 4 bytes of the window
	uint32_t crow;   // context `lane`: qe << 16 | switch << 13 | mps << 12 | nlps << 6 | nmps   (Table C.2 row + MPS)
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

// packed Table C.2 rows in shared memory (same format as MqD::crow, MPS bit clear)
struct DecTab { uint32_t row[47]; };

__device__ __forceinline__ uint32_t mqd_decode(MqD &q, const DecTab &T, uint32_t cx, int lane) {
	const uint32_t row = __shfl_sync(0xffffffffu, q.crow, cx);
	const uint32_t qs = row & 0xFFFF0000u, mps = (row >> 12) & 1u;
	q.a -= qs;
	bool lps;
	if (q.c < qs) { // (C >> 16) < Qe : the LPS sub-interval
		lps = q.a >= qs; // conditional exchange
		q.a = qs;
	} else {
		q.c -= qs;
		if (q.a & 0x80000000u) return mps;
		lps = q.a < qs;
	}
	const uint32_t next = lps ? (row >> 6) & 63u : row & 63u;
	const uint32_t nmps = lps ? mps ^ ((row >> 13) & 1u) : mps;
	const uint32_t nrow = T.row[next] | (nmps << 12);
	if (lane == (int) cx) q.crow = nrow;
	int sh = __clz(q.a);
	q.a <<= sh;
	if (sh <= q.ct) { q.c <<= sh; q.ct -= sh; }
	else {
		do { // RENORMD with BYTEIN whenever the bit counter runs out (mqc_dec_inl.h:136-147)
			if (q.ct == 0) mqd_bytein(q, lane);
			const int n = sh < q.ct ? sh : q.ct;
			q.c <<= n; q.ct -= n; sh -= n;
		} while (sh > 0);
	}
	return lps ? mps ^ 1u : mps;
}

// bits 4,7,10,13 (significance of the own column, rows 0..3) gathered into a nibble
__device__ __forceinline__ uint32_t own_sig4(uint32_t f) {
	return ((f >> 4) & 1u) | ((f >> 6) & 2u) | ((f >> 8) & 4u) | ((f >> 10) & 8u);
}
// rows whose 8-neighbourhood holds a significant sample
__device__ __forceinline__ uint32_t nbr4(uint32_t f) {
	return ((f & 0x1EFu) ? 1u : 0u) | ((f & (0x1EFu << 3)) ? 2u : 0u) | ((f & (0x1EFu << 6)) ? 4u : 0u) | ((f & (0x1EFu << 9)) ? 8u : 0u);
}

__global__ void __launch_bounds__(DEC_WARPS * 32, DEC_MIN_CTAS) t1_decode_kernel(const DecBlock *__restrict__ blocks,
		const DecInput *__restrict__ inputs, uint32_t nblocks, const uint8_t *__restrict__ data, uint32_t max_planes,
		uint64_t *__restrict__ plane_scratch) {
	__shared__ DecWarp warps[DEC_WARPS];
	__shared__ uint8_t Lzc[4][512]; // zero-coding context by the 9 neighbourhood bits of a stripe-column word
	__shared__ uint8_t Lsc[256];
	__shared__ DecTab T;
	for (int i = threadIdx.x; i < 47; i += blockDim.x) {
		const uint32_t r = c_mq[i];
		T.row[i] = (r << 16) | ((r >> 28) & 1u) << 13 | ((r >> 22) & 63u) << 6 | ((r >> 16) & 63u);
	}
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
		int o = i >> 9, n9 = i & 511;
		int idx8 = (n9 & 7) | ((n9 >> 3) & 1) << 3 | ((n9 >> 5) & 1) << 4 | ((n9 >> 6) & 7) << 5;
		Lzc[o][n9] = c_zc[o][idx8];
	}
	for (int i = threadIdx.x; i < 256; i += blockDim.x) Lsc[i] = c_sc[i];
	__syncthreads();

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t bid = blockIdx.x * DEC_WARPS + wid;
	if (bid >= nblocks) return;
	DecWarp &W = warps[wid];
	uint64_t *planes = plane_scratch + (size_t) bid * max_planes * 64; // [plane-1][row], global scratch
	const DecBlock B = blocks[bid];
	const DecInput I = inputs[bid];
	const int w = B.w, h = B.h;
	const uint8_t *zc = Lzc[B.orient];
	const int numbps = (int) I.numbps;
	const bool two = w > 32;
	const bool valid0 = lane < w, valid1 = lane + 32 < w;

	const bool empty = I.numpasses == 0 || I.data_len == 0 || numbps == 0 || numbps > (int) max_planes; // T1Part1.cpp:139
	if (empty) {
		// nothing decoded: the reference leaves the zero-initialised tile buffer untouched
		for (int y = 0; y < h; ++y)
			for (int x = lane; x < w; x += 32) B.dst[(size_t) y * B.stride + x] = 0;
		return;
	}
	for (int i = lane; i < 18 * 64; i += 32) (&W.F[0][0])[i] = 0;
	for (int i = lane; i < 64; i += 32) { W.pcur[i] = 0; W.lastc[i] = 0; }
	__syncwarp();

	MqD q;
	q.buf = data + I.data_offset;
	q.len = I.data_len;
	q.pos = 0;
	q.crow = T.row[lane == CTX_ZC0 ? 4 : lane == CTX_AGG ? 3 : lane == CTX_UNI ? 46 : 0]; // mqc_dec.cpp:207-214, mps = 0
	mqd_fill(q, 0, lane);
	q.c = mqd_byte(q, 0, lane) << 16; // INITDEC, mqc_dec.cpp:179-201
	mqd_bytein(q, lane);
	q.c <<= 7;
	q.ct -= 7;
	q.a = 0x80000000u;

	// plane of the last pass that will run: passes go cln(numbps), then sig/ref/cln per lower plane
	const int npass_eff = min((int) I.numpasses, 3 * numbps - 2);
	const int finalplane = numbps - (npass_eff + 1) / 3;
	const int nstripes = (h + 3) >> 2;

	int bp1 = numbps, type = 2;
	for (int pass = 0; pass < npass_eff; ++pass) {
		const bool track_last = bp1 == finalplane;
		for (int s = 0; s < nstripes; ++s) {
			const int nk = min(4, h - 4 * s);
			uint32_t f0 = W.F[s + 1][lane], f1 = two ? W.F[s + 1][lane + 32] : 0u;
			// which columns hold a sample this pass can code (ballot of a per-lane test)
			const uint32_t rows = (1u << nk) - 1u;
			// rows of a column this pass can code, from its word
			auto todo = [&](uint32_t f) -> uint32_t {
				const uint32_t sig = own_sig4(f), vis = (f >> 24) & 0xFu;
				if (type == 0) return ~sig & ~vis & nbr4(f) & rows;
				if (type == 1) return sig & ~vis & rows;
				return ~sig & ~vis & rows;
			};
			auto wants = [&](uint32_t f) -> bool { return todo(f) != 0; };
			uint64_t cols = (uint64_t) __ballot_sync(0xffffffffu, valid0 && wants(f0));
			if (two) cols |= (uint64_t) __ballot_sync(0xffffffffu, valid1 && wants(f1)) << 32;
			uint32_t mb0 = 0, mb1 = 0, lc0 = 0, lc1 = 0; // per-lane nibbles: magnitude bit decoded / sample coded
			while (cols) {
				const int x = __ffsll((long long) cols) - 1;
				cols &= cols - 1;
				const int src = x & 31;
				const bool hi = x >= 32;
				uint32_t f = __shfl_sync(0xffffffffu, hi ? f1 : f0, src);
				uint32_t mb = 0, lc = 0;
				uint32_t cand = todo(f);
				if (type == 1) {
					while (cand) {
						const int k = __ffs(cand) - 1;
						cand &= cand - 1;
						const uint32_t ctx = (f & (1u << (28 + k))) ? CTX_MR0 + 2 : ((f >> (3 * k)) & 0x1EFu) ? CTX_MR0 + 1 : CTX_MR0;
						if (mqd_decode(q, T, ctx, lane)) mb |= 1u << k;
					}
					const uint32_t done4 = todo(f);
					f |= done4 << 28;
					lc = done4;
				} else {
					int kimp = -1; // row whose 1 is implied by the run-length code
					if (type == 2 && nk == 4 && (f & 0x0F03FFFFu) == 0) { // run-length mode: nothing significant or visited around
						if (!mqd_decode(q, T, CTX_AGG, lane)) continue;
						kimp = (int) mqd_decode(q, T, CTX_UNI, lane) << 1;
						kimp |= (int) mqd_decode(q, T, CTX_UNI, lane);
						cand &= ~((1u << kimp) - 1u);
					}
					uint32_t fW = 0, fE = 0;
					bool have_nb = false;
					while (cand) {
						const int k = __ffs(cand) - 1;
						cand &= cand - 1;
						const uint32_t n9 = (f >> (3 * k)) & 0x1FFu;
						uint32_t d = 1;
						if (k != kimp) d = mqd_decode(q, T, zc[n9], lane);
						if (type == 0) f |= 1u << (24 + k);
						if (!d) continue;
						if (!have_nb) { // signs of the west / east columns live in their owners' words
							const int xm = x - 1, xp = x + 1;
							const uint32_t a0 = __shfl_sync(0xffffffffu, xm >= 32 ? f1 : f0, xm & 31);
							const uint32_t a1 = __shfl_sync(0xffffffffu, xp >= 32 ? f1 : f0, xp & 31);
							fW = xm >= 0 ? a0 : 0u;
							fE = xp < w ? a1 : 0u;
							have_nb = true;
						}
						const uint32_t sN = n9 >> 1 & 1, sW = n9 >> 3 & 1, sE = n9 >> 5 & 1, sS = n9 >> 7 & 1;
						const uint32_t gN = f >> (18 + k) & 1, gS = f >> (20 + k) & 1, gW = fW >> (19 + k) & 1, gE = fE >> (19 + k) & 1;
						const uint32_t idx = sN | sW << 1 | sE << 2 | sS << 3 | (gN & sN) << 4 | (gW & sW) << 5 | (gE & sE) << 6 | (gS & sS) << 7;
						const uint32_t v = Lsc[idx];
						const uint32_t neg = mqd_decode(q, T, v & 31u, lane) ^ (v >> 5);
						f |= fsig(k + 1, 1) | (neg << (19 + k));
						mb |= 1u << k;
						lc |= 1u << k;
						// in the significance pass the row below may have become codable through this sample
						if (type == 0 && k + 1 < nk && !(f & (fsig(k + 2, 1) | (1u << (25 + k))))) cand |= 2u << k;
						// the sample is the east neighbour of column x-1 and the west neighbour of column x+1
						if (x > 0 && lane == ((x - 1) & 31)) { if (x - 1 >= 32) f1 |= fsig(k + 1, 2); else f0 |= fsig(k + 1, 2); }
						if (x + 1 < w) {
							if (lane == ((x + 1) & 31)) { if (x + 1 >= 32) f1 |= fsig(k + 1, 0); else f0 |= fsig(k + 1, 0); }
							if (type == 0) cols |= 1ull << (x + 1); // may have become codable in this pass
						}
						// rows -1 / 4 of the stripes below / above
						if ((k == 0 && s > 0) || (k == 3 && s + 1 < nstripes)) {
							const int ts = k == 0 ? s : s + 2; // index into F (stripe + 1)
							const int r = k == 0 ? 5 : 0;
							if (lane < 3) {
								const int cc = x + lane - 1;
								if (cc >= 0 && cc < w)
									atomicOr(&W.F[ts][cc], fsig(r, 2 - lane) | (lane == 1 ? neg << (18 + r) : 0u));
							}
						}
					}
				}
				if (lane == src) {
					if (hi) { f1 = f; mb1 |= mb; lc1 |= lc; } else { f0 = f; mb0 |= mb; lc0 |= lc; }
				}
			}
			if (type == 2) { f0 &= ~F_PI_ALL; f1 &= ~F_PI_ALL; }
			W.F[s + 1][lane] = f0;
			if (two) W.F[s + 1][lane + 32] = f1;
			// per-lane nibbles -> row masks of the current plane
			if (__any_sync(0xffffffffu, (mb0 | mb1 | lc0 | lc1) != 0)) {
				#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint64_t m = (uint64_t) __ballot_sync(0xffffffffu, mb0 >> k & 1) | (uint64_t) __ballot_sync(0xffffffffu, mb1 >> k & 1) << 32;
					const uint64_t l = (uint64_t) __ballot_sync(0xffffffffu, lc0 >> k & 1) | (uint64_t) __ballot_sync(0xffffffffu, lc1 >> k & 1) << 32;
					if (lane == k && k < nk) {
						W.pcur[4 * s + k] |= m;
						if (track_last) W.lastc[4 * s + k] |= l;
					}
				}
			}
			__syncwarp();
		}
		if (++type == 3) {
			// plane finished: park its magnitude bits in the scratch area
			for (int i = lane; i < 64; i += 32) { planes[(size_t) (bp1 - 1) * 64 + i] = W.pcur[i]; W.pcur[i] = 0; }
			__syncwarp();
			type = 0;
			bp1--;
		}
	}
	const int lastplane = finalplane;
	if (type != 0) { // the last plane was left unfinished (no cleanup pass): park what there is
		for (int i = lane; i < 64; i += 32) planes[(size_t) (lastplane - 1) * 64 + i] = W.pcur[i];
	}
	__syncwarp();
	// ---- reconstruct, de-quantise, scatter (T1Part1.cpp:216-329) -----------------------------
	for (int y = 0; y < h; ++y) {
		const int s = y >> 2, k = y & 3;
		const uint64_t lrow = W.lastc[y];
		uint32_t mag0 = 0, mag1 = 0;
		for (int p = lastplane; p <= numbps; ++p) {
			const uint64_t m = planes[(size_t) (p - 1) * 64 + y];
			mag0 |= (uint32_t) (m >> lane & 1) << p;
			mag1 |= (uint32_t) (m >> (lane + 32) & 1) << p;
		}
		#pragma unroll
		for (int half = 0; half < 2; ++half) {
			const int x = lane + 32 * half;
			if (x >= w) continue;
			const uint32_t f = W.F[s + 1][x];
			int32_t v = 0;
			if (f & fsig(k + 1, 1)) {
				uint32_t mag = half ? mag1 : mag0;
				mag |= (lrow >> x & 1) ? (1u << lastplane) >> 1 : 1u << lastplane;
				v = (f >> (19 + k) & 1) ? -(int32_t) mag : (int32_t) mag;
			}
			int32_t o;
			if (B.reversible) o = v / 2;
			else o = __float_as_int(__fmul_rn((float) v, B.stepsize));
			B.dst[(size_t) y * B.stride + x] = o;
		}
	}
}

static bool g_dec_tables_ready = false;

size_t t1_decode_scratch_bytes(uint32_t nblocks, uint32_t max_planes) {
	return (size_t) nblocks * (max_planes < 1 ? 1 : max_planes) * 64 * sizeof(uint64_t);
}

void launch_t1_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data,
		uint32_t max_planes, uint64_t *plane_scratch, cudaStream_t s) {
	if (!nblocks) return;
	if (!g_dec_tables_ready) { build_and_upload_t1_tables(); g_dec_tables_ready = true; }
	if (max_planes < 1) max_planes = 1;
	t1_decode_kernel<<<(nblocks + DEC_WARPS - 1) / DEC_WARPS, DEC_WARPS * 32, 0, s>>>(blocks, inputs, nblocks, data, max_planes,
			plane_scratch);
}

} // namespace gb
