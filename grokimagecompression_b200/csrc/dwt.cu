// K2 / K3: forward and inverse wavelet transforms, one decomposition level per launch, every
// lifting step of both directions in one pass over the data.
//
// forward 5/3  WaveletForward.h:40-161 + dwt53.cpp:150-169   (int32, exact)
// forward 9/7  WaveletForward.h:40-161 + dwt97.cpp:90-123    (int32 13-bit fixed point, exact)
// inverse 5/3  dwt.cpp:724-858, 256-363, 661-718             (int32, exact)
// inverse 9/7  dwt.cpp:1544-1738, 1413-1537, constants 172-178 (fp32, multiply-then-add, no FMA)
//
// A CTA owns a TW x TH tile of the interleaved (spatial) domain of one plane, stages it in shared
// memory with a halo of 2 (5/3) or 4 (9/7) samples, extended by whole-sample symmetric reflection
// at the region border, runs the vertical and the horizontal lifting there and writes the four
// sub-bands (forward) or the reconstructed tile (inverse) once: HBM sees one read and one write of
// the level's region (8 B/sample algorithmic).  Launches are out of place (ping-pong planes), so
// no CTA ever reads what another CTA of the same launch writes.
//
// Lifting on a symmetrically extended signal keeps the signal symmetric, so it produces exactly
// the reference's clamped-neighbour results at the borders (the (c+c)*x special case of
// dwt.cpp:1463-1469 is the same rounding as (x+x)*c).  Lines of length 1 are the exception and are
// handled explicitly (dwt53.cpp:160, dwt.cpp:344-349, 1482-1490).
#include "common.cuh"

namespace gb {

constexpr int TW = 64, TH = 64;

void dwt_tile_shape(uint32_t *tw, uint32_t *th) { *tw = TW; *th = TH; }

__device__ __forceinline__ int reflect(int i, int len) {
	// whole-sample symmetric extension, any distance
	if (len == 1) return 0;
	int p = 2 * (len - 1);
	i %= p;
	if (i < 0) i += p;
	return i >= len ? p - i : i;
}

__device__ __forceinline__ int32_t fix13(int32_t a, int32_t b) {
	return (int32_t) (((int64_t) a * (int64_t) b + 4096) >> 13);
}

__device__ __forceinline__ const DwtPlane &find_plane(const DwtPlane *planes, uint32_t &cta) {
	// planes are sorted by first_cta; linear walk is fine for the few hundred planes of a launch
	uint32_t lo = 0;
	while (planes[lo].first_cta + planes[lo].tiles_x * planes[lo].tiles_y <= cta) ++lo;
	cta -= planes[lo].first_cta;
	return planes[lo];
}

// One lifting step over the shared tile.  Updates samples whose LOCAL index along the lifted axis
// has parity `par`, using both neighbours; the outermost line of the tile has no neighbour and is
// left alone (it is halo).  VERT: axis = rows.  Threads walk the non-lifted axis fastest for the
// vertical pass (conflict-free rows) and the lifted axis slowest for the horizontal pass with an
// odd pitch (conflict-free columns).
template<int OP, bool VERT, int PITCH, typename T>
__device__ __forceinline__ void lift(T *sm, int nrows, int ncols, int par, int k) {
	int nlift = VERT ? nrows : ncols;
	int first = par ? 1 : 2;              // skip local index 0
	int cnt = (nlift - first) / 2;         // indices first, first+2, ... <= nlift-2
	if (cnt < 0) cnt = 0;
	int other = VERT ? ncols : nrows;
	for (int idx = threadIdx.x; idx < cnt * other; idx += blockDim.x) {
		int o = idx % other, l = first + 2 * (idx / other);
		T *p = VERT ? &sm[l * PITCH + o] : &sm[o * PITCH + l];
		const int step = VERT ? PITCH : 1;
		if (OP == 0) *p -= (p[-step] + p[step]) >> 1;                 // 5/3 predict
		else if (OP == 1) *p += (p[-step] + p[step] + 2) >> 2;        // 5/3 update
		else if (OP == 2) *p -= (p[-step] + p[step] + 2) >> 2;        // 5/3 inverse update
		else if (OP == 3) *p += (p[-step] + p[step]) >> 1;            // 5/3 inverse predict
		else if (OP == 4) *p -= fix13(p[-step] + p[step], k);         // 9/7 analysis
		else if (OP == 5) *p += fix13(p[-step] + p[step], k);
	}
	__syncthreads();
}

template<bool VERT, int PITCH>
__device__ __forceinline__ void scale97_fwd(int32_t *sm, int nrows, int ncols, int low_par) {
	for (int idx = threadIdx.x; idx < nrows * ncols; idx += blockDim.x) {
		int r = idx / ncols, c = idx % ncols;
		int l = VERT ? r : c;
		int32_t *p = &sm[r * PITCH + c];
		*p = fix13(*p, (l & 1) == low_par ? 6659 : 5039);
	}
	__syncthreads();
}

template<bool VERT, int PITCH>
__device__ __forceinline__ void lift97f(float *sm, int nrows, int ncols, int par, float c) {
	int nlift = VERT ? nrows : ncols;
	int first = par ? 1 : 2;
	int cnt = (nlift - first) / 2;
	if (cnt < 0) cnt = 0;
	int other = VERT ? ncols : nrows;
	for (int idx = threadIdx.x; idx < cnt * other; idx += blockDim.x) {
		int o = idx % other, l = first + 2 * (idx / other);
		float *p = VERT ? &sm[l * PITCH + o] : &sm[o * PITCH + l];
		const int step = VERT ? PITCH : 1;
		*p = __fadd_rn(*p, __fmul_rn(__fadd_rn(p[-step], p[step]), c));
	}
	__syncthreads();
}

template<bool VERT, int PITCH>
__device__ __forceinline__ void scale97_inv(float *sm, int nrows, int ncols, int low_par) {
	for (int idx = threadIdx.x; idx < nrows * ncols; idx += blockDim.x) {
		int r = idx / ncols, c = idx % ncols;
		int l = VERT ? r : c;
		float *p = &sm[r * PITCH + c];
		*p = __fmul_rn(*p, (l & 1) == low_par ? 1.230174105f : 1.625732422f);
	}
	__syncthreads();
}

template<bool REV>
__global__ void __launch_bounds__(256) dwt_fwd_kernel(const DwtPlane *__restrict__ planes) {
	constexpr int H = REV ? 2 : 4;
	constexpr int NR = TH + 2 * H, NC = TW + 2 * H, PITCH = NC | 1;
	__shared__ int32_t sm[NR * PITCH];

	uint32_t cta = blockIdx.x;
	const DwtPlane &P = find_plane(planes, cta);
	const int X0 = (cta % P.tiles_x) * TW, Y0 = (cta / P.tiles_x) * TH;
	const int rw = P.rw, rh = P.rh;

	// stage tile + halo, reflected at the region border
	for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x) {
		int r = idx / NC, c = idx % NC;
		int gy = reflect(Y0 + r - H, rh), gx = reflect(X0 + c - H, rw);
		sm[r * PITCH + c] = P.src[(size_t) gy * P.src_stride + gx];
	}
	__syncthreads();

	// local index l <-> region index X0 + l - H ; H and X0 are even, so parity(l) == parity(region index).
	// A sample is high-pass when (region index + cas) is odd.
	const int hpx = 1 - (int) P.cas_x, hpy = 1 - (int) P.cas_y; // local parity of high-pass samples
	// vertical first (WaveletForward.h:91-121), then horizontal (:124-152)
	if (rh > 1) {
		if (REV) {
			lift<0, true, PITCH>(sm, NR, NC, hpy, 0);
			lift<1, true, PITCH>(sm, NR, NC, 1 - hpy, 0);
		} else {
			lift<4, true, PITCH>(sm, NR, NC, hpy, 12994);
			lift<4, true, PITCH>(sm, NR, NC, 1 - hpy, 434);
			lift<5, true, PITCH>(sm, NR, NC, hpy, 7233);
			lift<5, true, PITCH>(sm, NR, NC, 1 - hpy, 3633);
			scale97_fwd<true, PITCH>(sm, NR, NC, 1 - hpy);
		}
	} else if (REV && P.cas_y) { // single high-pass row: doubled (dwt53.cpp:160)
		for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x)
			sm[(idx / NC) * PITCH + idx % NC] *= 2;
		__syncthreads();
	}
	if (rw > 1) {
		if (REV) {
			lift<0, false, PITCH>(sm, NR, NC, hpx, 0);
			lift<1, false, PITCH>(sm, NR, NC, 1 - hpx, 0);
		} else {
			lift<4, false, PITCH>(sm, NR, NC, hpx, 12994);
			lift<4, false, PITCH>(sm, NR, NC, 1 - hpx, 434);
			lift<5, false, PITCH>(sm, NR, NC, hpx, 7233);
			lift<5, false, PITCH>(sm, NR, NC, 1 - hpx, 3633);
			scale97_fwd<false, PITCH>(sm, NR, NC, 1 - hpx);
		}
	} else if (REV && P.cas_x) {
		for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x)
			sm[(idx / NC) * PITCH + idx % NC] *= 2;
		__syncthreads();
	}

	// de-interleave to Mallat layout (dwt_utils.cpp:84-127): low halves first.
	// Thread order: sub-band quadrant, row, column -> 128-byte row segments per quadrant.
	constexpr int QW = TW / 2, QH = TH / 2;
	for (int idx = threadIdx.x; idx < TW * TH; idx += blockDim.x) {
		int q = idx / (QW * QH), rem = idx % (QW * QH);
		int qr = rem / QW, qc = rem % QW;
		int hx = q & 1, hy = q >> 1; // 1: high-pass in that direction
		// local (unhaloed) coordinate with the wanted parity
		int lx = 2 * qc + (hx ? hpx : 1 - hpx), ly = 2 * qr + (hy ? hpy : 1 - hpy);
		int gx = X0 + lx, gy = Y0 + ly;
		if (gx >= rw || gy >= rh) continue;
		int ox = (gx >> 1) + (hx ? (int) P.sw : 0), oy = (gy >> 1) + (hy ? (int) P.sh : 0);
		P.dst[(size_t) oy * P.dst_stride + ox] = sm[(ly + H) * PITCH + lx + H];
	}
}

template<bool REV>
__global__ void __launch_bounds__(256) dwt_inv_kernel(const DwtPlane *__restrict__ planes) {
	constexpr int H = REV ? 2 : 4;
	constexpr int NR = TH + 2 * H, NC = TW + 2 * H, PITCH = NC | 1;
	__shared__ int32_t sm[NR * PITCH];

	uint32_t cta = blockIdx.x;
	const DwtPlane &P = find_plane(planes, cta);
	const int X0 = (cta % P.tiles_x) * TW, Y0 = (cta / P.tiles_x) * TH;
	const int rw = P.rw, rh = P.rh;
	const int hpx = 1 - (int) P.cas_x, hpy = 1 - (int) P.cas_y;

	// interleave the four sub-bands into the tile (dwt.cpp:1219-1316 for 9/7; implicit in the
	// single-sweep 5/3 routines).  LL comes from the previous level's output, the rest from the
	// coefficient plane.
	for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x) {
		int r = idx / NC, c = idx % NC;
		int gy = reflect(Y0 + r - H, rh), gx = reflect(X0 + c - H, rw);
		int hx = (gx & 1) == hpx, hy = (gy & 1) == hpy;
		int sx = (gx >> 1) + (hx ? (int) P.sw : 0), sy = (gy >> 1) + (hy ? (int) P.sh : 0);
		int32_t v;
		if (!hx && !hy) v = P.src[(size_t) sy * P.src_stride + sx];
		else v = P.band[(size_t) sy * P.band_stride + sx];
		sm[r * PITCH + c] = v;
	}
	__syncthreads();

	// horizontal first, then vertical (dwt.cpp:775-853, 1586-1733)
	if (REV) {
		if (rw > 1) {
			lift<2, false, PITCH>(sm, NR, NC, 1 - hpx, 0);
			lift<3, false, PITCH>(sm, NR, NC, hpx, 0);
		} else if (P.cas_x) {
			for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x)
				sm[(idx / NC) * PITCH + idx % NC] /= 2; // C division (dwt.cpp:349)
			__syncthreads();
		}
		if (rh > 1) {
			lift<2, true, PITCH>(sm, NR, NC, 1 - hpy, 0);
			lift<3, true, PITCH>(sm, NR, NC, hpy, 0);
		} else if (P.cas_y) {
			for (int idx = threadIdx.x; idx < NR * NC; idx += blockDim.x)
				sm[(idx / NC) * PITCH + idx % NC] /= 2;
			__syncthreads();
		}
	} else {
		float *fs = reinterpret_cast<float*>(sm);
		if (rw > 1) {
			scale97_inv<false, PITCH>(fs, NR, NC, 1 - hpx);
			lift97f<false, PITCH>(fs, NR, NC, 1 - hpx, -0.443506852f);
			lift97f<false, PITCH>(fs, NR, NC, hpx, -0.882911075f);
			lift97f<false, PITCH>(fs, NR, NC, 1 - hpx, 0.052980118f);
			lift97f<false, PITCH>(fs, NR, NC, hpx, 1.586134342f);
		}
		if (rh > 1) {
			scale97_inv<true, PITCH>(fs, NR, NC, 1 - hpy);
			lift97f<true, PITCH>(fs, NR, NC, 1 - hpy, -0.443506852f);
			lift97f<true, PITCH>(fs, NR, NC, hpy, -0.882911075f);
			lift97f<true, PITCH>(fs, NR, NC, 1 - hpy, 0.052980118f);
			lift97f<true, PITCH>(fs, NR, NC, hpy, 1.586134342f);
		}
	}

	for (int idx = threadIdx.x; idx < TW * TH; idx += blockDim.x) {
		int ly = idx / TW, lx = idx % TW;
		int gx = X0 + lx, gy = Y0 + ly;
		if (gx >= rw || gy >= rh) continue;
		P.dst[(size_t) gy * P.dst_stride + gx] = sm[(ly + H) * PITCH + lx + H];
	}
}

void launch_dwt_fwd(const DwtPlane *planes_dev, uint32_t total_ctas, int reversible, cudaStream_t s) {
	if (!total_ctas) return;
	if (reversible) dwt_fwd_kernel<true><<<total_ctas, 256, 0, s>>>(planes_dev);
	else dwt_fwd_kernel<false><<<total_ctas, 256, 0, s>>>(planes_dev);
}

void launch_dwt_inv(const DwtPlane *planes_dev, uint32_t total_ctas, int reversible, cudaStream_t s) {
	if (!total_ctas) return;
	if (reversible) dwt_inv_kernel<true><<<total_ctas, 256, 0, s>>>(planes_dev);
	else dwt_inv_kernel<false><<<total_ctas, 256, 0, s>>>(planes_dev);
}

} // namespace gb
