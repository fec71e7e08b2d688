// K1 / K1': DC level shift fused with the multi-component transforms.
//
// forward : TileProcessor::dc_level_shift_encode (TileProcessor.cpp:1449-1471) then
//           mct::encode_rev (mct.cpp:125-135) or mct::encode_irrev (mct.cpp:336-346)
// inverse : mct::decode_rev (mct.cpp:180-190) or mct::decode_irrev (mct.cpp:394-404) then
//           TileProcessor::dc_level_shift_decode (TileProcessor.cpp:1377-1432)
//
// HBM-bound element-wise kernels: every thread moves 16-byte vectors, one read and one write per
// sample (8 B/sample algorithmic), grid sized to a multiple of the SM count.
#include "common.cuh"

namespace gb {

__device__ __forceinline__ int32_t fix13(int32_t a, int32_t b) {
	return (int32_t) (((int64_t) a * (int64_t) b + 4096) >> 13);
}

struct Shift3 { int32_t s[3], lo[3], hi[3]; };

template<bool REV, bool SHIFT>
__device__ __forceinline__ void mct_fwd_px(int32_t &r, int32_t &g, int32_t &b, const Shift3 &p) {
	if (SHIFT) {
		if (REV) { r -= p.s[0]; g -= p.s[1]; b -= p.s[2]; }
		else { r = (r - p.s[0]) * 2048; g = (g - p.s[1]) * 2048; b = (b - p.s[2]) * 2048; }
	}
	if (REV) {
		int32_t y = (r + 2 * g + b) >> 2, u = b - g, v = r - g;
		r = y; g = u; b = v;
	} else {
		int32_t y = fix13(r, 2449) + fix13(g, 4809) + fix13(b, 934);
		int32_t u = -fix13(r, 1382) - fix13(g, 2714) + fix13(b, 4096);
		int32_t v = fix13(r, 4096) - fix13(g, 3430) - fix13(b, 666);
		r = y; g = u; b = v;
	}
}

__device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return min(max(v, lo), hi); }

template<bool REV, bool SHIFT>
__device__ __forceinline__ void mct_inv_px(int32_t &c0, int32_t &c1, int32_t &c2, const Shift3 &p) {
	if (REV) {
		int32_t y = c0, u = c1, v = c2;
		int32_t g = y - ((u + v) >> 2);
		c0 = v + g; c1 = g; c2 = u + g;
		if (SHIFT) {
			c0 = clampi(c0 + p.s[0], p.lo[0], p.hi[0]);
			c1 = clampi(c1 + p.s[1], p.lo[1], p.hi[1]);
			c2 = clampi(c2 + p.s[2], p.lo[2], p.hi[2]);
		}
	} else {
		float y = __int_as_float(c0), u = __int_as_float(c1), v = __int_as_float(c2);
		// multiply, then add: the reference's SSE path never fuses (mct.cpp:394-404)
		float r = __fadd_rn(y, __fmul_rn(v, 1.402f));
		float g = __fsub_rn(__fsub_rn(y, __fmul_rn(u, 0.34413f)), __fmul_rn(v, 0.71414f));
		float b = __fadd_rn(y, __fmul_rn(u, 1.772f));
		if (SHIFT) { // lrintf: round half to even, then shift and clamp
			c0 = clampi(__float2int_rn(r) + p.s[0], p.lo[0], p.hi[0]);
			c1 = clampi(__float2int_rn(g) + p.s[1], p.lo[1], p.hi[1]);
			c2 = clampi(__float2int_rn(b) + p.s[2], p.lo[2], p.hi[2]);
		} else {
			c0 = __float_as_int(r); c1 = __float_as_int(g); c2 = __float_as_int(b);
		}
	}
}

template<bool FWD, bool REV, bool SHIFT>
__global__ void __launch_bounds__(256) mct3_kernel(int32_t *__restrict__ c0, int32_t *__restrict__ c1,
		int32_t *__restrict__ c2, uint64_t n, Shift3 p) {
	uint64_t nvec = n >> 2;
	uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t nthr = (uint64_t) gridDim.x * blockDim.x;
	int4 *v0 = reinterpret_cast<int4*>(c0), *v1 = reinterpret_cast<int4*>(c1), *v2 = reinterpret_cast<int4*>(c2);
	for (uint64_t i = tid; i < nvec; i += nthr) {
		int4 a = v0[i], b = v1[i], c = v2[i];
		if (FWD) {
			mct_fwd_px<REV, SHIFT>(a.x, b.x, c.x, p); mct_fwd_px<REV, SHIFT>(a.y, b.y, c.y, p);
			mct_fwd_px<REV, SHIFT>(a.z, b.z, c.z, p); mct_fwd_px<REV, SHIFT>(a.w, b.w, c.w, p);
		} else {
			mct_inv_px<REV, SHIFT>(a.x, b.x, c.x, p); mct_inv_px<REV, SHIFT>(a.y, b.y, c.y, p);
			mct_inv_px<REV, SHIFT>(a.z, b.z, c.z, p); mct_inv_px<REV, SHIFT>(a.w, b.w, c.w, p);
		}
		v0[i] = a; v1[i] = b; v2[i] = c;
	}
	for (uint64_t i = (nvec << 2) + tid; i < n; i += nthr) {
		int32_t a = c0[i], b = c1[i], c = c2[i];
		if (FWD) mct_fwd_px<REV, SHIFT>(a, b, c, p); else mct_inv_px<REV, SHIFT>(a, b, c, p);
		c0[i] = a; c1[i] = b; c2[i] = c;
	}
}

template<bool FWD, bool REV>
__global__ void __launch_bounds__(256) dcshift_kernel(int32_t *__restrict__ x, uint64_t n, int32_t shift, int32_t lo,
		int32_t hi) {
	uint64_t nvec = n >> 2;
	uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t nthr = (uint64_t) gridDim.x * blockDim.x;
	auto f = [&](int32_t v) -> int32_t {
		if (FWD) return REV ? v - shift : (v - shift) * 2048;
		int32_t t = REV ? v : __float2int_rn(__int_as_float(v));
		return clampi(t + shift, lo, hi);
	};
	int4 *xv = reinterpret_cast<int4*>(x);
	for (uint64_t i = tid; i < nvec; i += nthr) {
		int4 a = xv[i];
		a.x = f(a.x); a.y = f(a.y); a.z = f(a.z); a.w = f(a.w);
		xv[i] = a;
	}
	for (uint64_t i = (nvec << 2) + tid; i < n; i += nthr)
		x[i] = f(x[i]);
}

static inline unsigned grid_for(uint64_t n) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	uint64_t want = (n / 4 + 255) / 256;
	uint64_t cap = (uint64_t) sms * 8;
	if (want < 1) want = 1;
	return (unsigned) (want < cap ? want : cap);
}

// all planes handed to these launchers are 256-byte aligned device allocations (api.cu)

void launch_dcshift_fwd(int32_t *x, uint64_t n, int32_t shift, int reversible, cudaStream_t s) {
	if (!n) return;
	if (reversible) dcshift_kernel<true, true><<<grid_for(n), 256, 0, s>>>(x, n, shift, 0, 0);
	else dcshift_kernel<true, false><<<grid_for(n), 256, 0, s>>>(x, n, shift, 0, 0);
}

void launch_dcshift_inv(int32_t *x, uint64_t n, int32_t shift, int reversible, int32_t lo, int32_t hi, cudaStream_t s) {
	if (!n) return;
	if (reversible) dcshift_kernel<false, true><<<grid_for(n), 256, 0, s>>>(x, n, shift, lo, hi);
	else dcshift_kernel<false, false><<<grid_for(n), 256, 0, s>>>(x, n, shift, lo, hi);
}

void launch_mct_fwd(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n, int32_t s0, int32_t s1, int32_t s2,
		int reversible, int do_shift, cudaStream_t s) {
	if (!n) return;
	Shift3 p = {{s0, s1, s2}, {0, 0, 0}, {0, 0, 0}};
	unsigned g = grid_for(n);
	if (reversible) {
		if (do_shift) mct3_kernel<true, true, true><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
		else mct3_kernel<true, true, false><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
	} else {
		if (do_shift) mct3_kernel<true, false, true><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
		else mct3_kernel<true, false, false><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
	}
}

void launch_mct_inv(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n, const int32_t shift[3], const int32_t lo[3],
		const int32_t hi[3], int reversible, int do_shift_clamp, cudaStream_t s) {
	if (!n) return;
	Shift3 p;
	for (int i = 0; i < 3; ++i) { p.s[i] = shift ? shift[i] : 0; p.lo[i] = lo ? lo[i] : 0; p.hi[i] = hi ? hi[i] : 0; }
	unsigned g = grid_for(n);
	if (reversible) {
		if (do_shift_clamp) mct3_kernel<false, true, true><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
		else mct3_kernel<false, true, false><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
	} else {
		if (do_shift_clamp) mct3_kernel<false, false, true><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
		else mct3_kernel<false, false, false><<<g, 256, 0, s>>>(c0, c1, c2, n, p);
	}
}

} // namespace gb
