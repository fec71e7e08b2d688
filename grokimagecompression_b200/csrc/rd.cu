// K7: PCRD preparation on the device (first half): the feasible truncation points of every code
// block and their distortion-rate slopes, RateControl::convexHull + slopeToLog (t2/RateControl.cpp:31-118, 159-168), from
// the pass tables the Tier-1 encoder has just left in HBM.  One thread per code block: the walk over a block's passes is
// a short serial recurrence (at most 3 * 31 - 2 passes), the blocks are independent.  The threshold bisection and the
// packet simulation that use the slopes stay in the host (TileProcessor.cpp:281-366, t2/T2.cpp).
//
// Arithmetic: fp64 with explicit round-to-nearest operations in the reference's order.  The one library call is log():
// CUDA's and glibc's are both within 1 ulp, and the result is truncated to 16 bits after scaling by 256 / ln 2, so the two can
// differ only when ln(slope) * 369.3 lies within ~1e-11 of an integer; the parity tests compare thousands of passes exactly.
#include "common.cuh"
#include <cmath>

namespace gb {

struct SlopeConsts { double cutoff, scale, cutoff_term, shift; };

__device__ __forceinline__ uint16_t slope_to_log(double slope, const SlopeConsts &k) {
	if (slope > k.cutoff) slope = k.cutoff;
	double v = __dadd_rn(__dsub_rn(__dmul_rn(log(slope), k.scale), k.cutoff_term), k.shift);
	if (v < 1) v = 1;
	if (v > 65535.0) v = 65535.0;
	return (uint16_t) v;
}

__global__ void __launch_bounds__(128) rd_slopes_kernel(const EncBlock *__restrict__ blocks, const EncResult *__restrict__ results,
		uint32_t nblocks, const uint32_t *__restrict__ rates, const double *__restrict__ dists, uint16_t *__restrict__ slopes,
		double *__restrict__ cache, SlopeConsts k) {
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks) return;
	// an overflowed block carries numpasses = 0xFFFFFFFF (t1_enc.cu): nothing to do, and never walk past the block's slots
	const uint32_t off = blocks[b].pass_offset, np = results[b].numpasses;
	const uint32_t n = np == 0xFFFFFFFFu ? 0u : min(np, blocks[b].max_passes);
	const uint32_t *rate = rates + off;  // cumulative bytes; the pass length is the difference (t1.cpp:1303-1324)
	const double *dist = dists + off;    // cumulative distortion decrease
	uint16_t *slope = slopes + off;
	double *sc = cache + off;
	for (uint32_t p = 0; p < n; ++p) {
		double dd = 0, dr = 0;
		int q = (int) p;
		slope[p] = 0;
		for (;;) {
			dr = __dadd_rn(dr, (double) (rate[q] - (q ? rate[q - 1] : 0u)));
			dd = __dadd_rn(dd, q == 0 ? dist[q] : __dsub_rn(dist[q], dist[q - 1]));
			if (dd <= 0) { slope[p] = 0; break; }
			--q;
			if (q == -1) { sc[p] = __ddiv_rn(dd, dr); slope[p] = slope_to_log(sc[p], k); break; }
			if (slope[q] == 0) continue;
			if (dr == 0) slope[q] = 0;
			else if (__dmul_rn(sc[q], dr) <= dd) slope[q] = 0;
			else {
				sc[p] = __ddiv_rn(dd, dr);
				slope[p] = slope_to_log(sc[p], k);
				if (slope[p] >= slope[q]) slope[q] = 0;
				break;
			}
		}
	}
}

void launch_rd_slopes(const EncBlock *blocks, const EncResult *results, uint32_t nblocks, const uint32_t *rates, const double *dists,
		uint16_t *slopes, double *cache, cudaStream_t s) {
	if (!nblocks) return;
	SlopeConsts k;
	k.cutoff = std::pow(2.0, 64);
	k.scale = 256 / std::log(2.0);
	k.cutoff_term = std::log(k.cutoff) * k.scale;
	k.shift = (double) (1 << 16);
	rd_slopes_kernel<<<(nblocks + 127) / 128, 128, 0, s>>>(blocks, results, nblocks, rates, dists, slopes, cache, k);
}

} // namespace gb
