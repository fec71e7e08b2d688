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

// ---- narrow-sample boundary (TileProcessor.cpp:1201-1258 copy-in, 1691-1921 copy-out) ---------------------------
// The host's images hold 8 / 12 / 16-bit samples; the reference expands them to int32 on the host before the tile coder
// sees them.  Here the packed samples cross PCIe as they are and the widening (forward) / clamp + narrowing (inverse) is
// fused into the level-shift + MCT pass: T = uint8_t / int8_t / uint16_t / int16_t.  Algorithmic traffic: sizeof(T) + 4
// bytes per sample.
// Work unit = 4 samples: one 4- or 8-byte packed load / store and one 16-byte int32 store / load per plane, every warp
// access a run of consecutive bytes; four units per thread and trip keep the loads in flight.
template<typename T>
__device__ __forceinline__ void load4(const T *p, uint64_t u, int32_t (&out)[4]) {
	if (sizeof(T) == 1) {
		const uint32_t w = reinterpret_cast<const uint32_t*>(p)[u];
		#pragma unroll
		for (int k = 0; k < 4; ++k) out[k] = (int32_t) (T) ((w >> (8 * k)) & 0xFFu);
	} else {
		const uint2 w = reinterpret_cast<const uint2*>(p)[u];
		out[0] = (int32_t) (T) (w.x & 0xFFFFu); out[1] = (int32_t) (T) (w.x >> 16);
		out[2] = (int32_t) (T) (w.y & 0xFFFFu); out[3] = (int32_t) (T) (w.y >> 16);
	}
}

template<typename T>
__device__ __forceinline__ void store4(T *p, uint64_t u, const int32_t (&in)[4]) {
	if (sizeof(T) == 1)
		reinterpret_cast<uint32_t*>(p)[u] = ((uint32_t) in[0] & 0xFFu) | ((uint32_t) in[1] & 0xFFu) << 8 | ((uint32_t) in[2] & 0xFFu) << 16 | (uint32_t) in[3] << 24;
	else
		reinterpret_cast<uint2*>(p)[u] = make_uint2(((uint32_t) in[0] & 0xFFFFu) | (uint32_t) in[1] << 16, ((uint32_t) in[2] & 0xFFFFu) | (uint32_t) in[3] << 16);
}

constexpr int PK_UNROLL = 4;

// forward: packed src planes -> int32 planes (level shift, x2048 for 9/7, RCT / ICT)
template<bool REV, typename T>
__global__ void __launch_bounds__(256) mct3_fwd_packed_kernel(const T *__restrict__ s0, const T *__restrict__ s1, const T *__restrict__ s2,
		int32_t *__restrict__ c0, int32_t *__restrict__ c1, int32_t *__restrict__ c2, uint64_t n, Shift3 p) {
	const uint64_t nunits = n >> 2;
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t u0 = tid; u0 < nunits; u0 += nthr * PK_UNROLL) {
		int32_t a[PK_UNROLL][4], b[PK_UNROLL][4], c[PK_UNROLL][4];
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u < nunits) { load4<T>(s0, u, a[j]); load4<T>(s1, u, b[j]); load4<T>(s2, u, c[j]); }
		}
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u >= nunits) break;
			#pragma unroll
			for (int k = 0; k < 4; ++k) mct_fwd_px<REV, true>(a[j][k], b[j][k], c[j][k], p);
			reinterpret_cast<int4*>(c0)[u] = make_int4(a[j][0], a[j][1], a[j][2], a[j][3]);
			reinterpret_cast<int4*>(c1)[u] = make_int4(b[j][0], b[j][1], b[j][2], b[j][3]);
			reinterpret_cast<int4*>(c2)[u] = make_int4(c[j][0], c[j][1], c[j][2], c[j][3]);
		}
	}
	for (uint64_t i = (nunits << 2) + tid; i < n; i += nthr) {
		int32_t a = (int32_t) s0[i], b = (int32_t) s1[i], c = (int32_t) s2[i];
		mct_fwd_px<REV, true>(a, b, c, p);
		c0[i] = a; c1[i] = b; c2[i] = c;
	}
}

// inverse: int32 / fp32 planes -> packed dst planes (inverse RCT / ICT, rounding, level shift, clamp, narrowing)
template<bool REV, typename T>
__global__ void __launch_bounds__(256) mct3_inv_packed_kernel(const int32_t *__restrict__ c0, const int32_t *__restrict__ c1,
		const int32_t *__restrict__ c2, T *__restrict__ d0, T *__restrict__ d1, T *__restrict__ d2, uint64_t n, Shift3 p) {
	const uint64_t nunits = n >> 2;
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t u0 = tid; u0 < nunits; u0 += nthr * PK_UNROLL) {
		int4 va[PK_UNROLL], vb[PK_UNROLL], vc[PK_UNROLL];
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u < nunits) { va[j] = reinterpret_cast<const int4*>(c0)[u]; vb[j] = reinterpret_cast<const int4*>(c1)[u]; vc[j] = reinterpret_cast<const int4*>(c2)[u]; }
		}
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u >= nunits) break;
			int32_t a[4] = {va[j].x, va[j].y, va[j].z, va[j].w}, b[4] = {vb[j].x, vb[j].y, vb[j].z, vb[j].w}, c[4] = {vc[j].x, vc[j].y, vc[j].z, vc[j].w};
			#pragma unroll
			for (int k = 0; k < 4; ++k) mct_inv_px<REV, true>(a[k], b[k], c[k], p);
			store4<T>(d0, u, a); store4<T>(d1, u, b); store4<T>(d2, u, c);
		}
	}
	for (uint64_t i = (nunits << 2) + tid; i < n; i += nthr) {
		int32_t a = c0[i], b = c1[i], c = c2[i];
		mct_inv_px<REV, true>(a, b, c, p);
		d0[i] = (T) a; d1[i] = (T) b; d2[i] = (T) c;
	}
}

// one component without MCT: forward widen + shift, inverse round + shift + clamp + narrow
template<bool FWD, bool REV, typename T>
__global__ void __launch_bounds__(256) dcshift_packed_kernel(const void *__restrict__ src, void *__restrict__ dst, uint64_t n, int32_t shift,
		int32_t lo, int32_t hi) {
	const uint64_t nunits = n >> 2;
	const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t) gridDim.x * blockDim.x;
	auto f = [&](int32_t v) -> int32_t {
		if (FWD) return REV ? v - shift : (v - shift) * 2048;
		const int32_t t = REV ? v : __float2int_rn(__int_as_float(v));
		return clampi(t + shift, lo, hi);
	};
	for (uint64_t u0 = tid; u0 < nunits; u0 += nthr * PK_UNROLL) {
		int32_t a[PK_UNROLL][4];
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u >= nunits) continue;
			if (FWD) load4<T>(static_cast<const T*>(src), u, a[j]);
			else { const int4 v = reinterpret_cast<const int4*>(src)[u]; a[j][0] = v.x; a[j][1] = v.y; a[j][2] = v.z; a[j][3] = v.w; }
		}
		#pragma unroll
		for (int j = 0; j < PK_UNROLL; ++j) {
			const uint64_t u = u0 + (uint64_t) j * nthr;
			if (u >= nunits) break;
			#pragma unroll
			for (int k = 0; k < 4; ++k) a[j][k] = f(a[j][k]);
			if (FWD) reinterpret_cast<int4*>(dst)[u] = make_int4(a[j][0], a[j][1], a[j][2], a[j][3]);
			else store4<T>(static_cast<T*>(dst), u, a[j]);
		}
	}
	for (uint64_t i = (nunits << 2) + tid; i < n; i += nthr) {
		if (FWD) static_cast<int32_t*>(dst)[i] = f((int32_t) static_cast<const T*>(src)[i]);
		else static_cast<T*>(dst)[i] = (T) f(static_cast<const int32_t*>(src)[i]);
	}
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

// packed variants: sample_bytes 1 or 2, sgnd selects int8 / int16
template<typename F>
static void with_sample_type(uint32_t sample_bytes, int sgnd, F &&f) {
	if (sample_bytes == 1) { if (sgnd) f((int8_t) 0); else f((uint8_t) 0); }
	else { if (sgnd) f((int16_t) 0); else f((uint16_t) 0); }
}

void launch_mct_fwd_packed(const void *s0, const void *s1, const void *s2, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n,
		int32_t sh0, int32_t sh1, int32_t sh2, int reversible, uint32_t sample_bytes, int sgnd, cudaStream_t s) {
	if (!n) return;
	Shift3 p = {{sh0, sh1, sh2}, {0, 0, 0}, {0, 0, 0}};
	const unsigned g = grid_for(n / PK_UNROLL);
	with_sample_type(sample_bytes, sgnd, [&](auto t) {
		using T = decltype(t);
		if (reversible) mct3_fwd_packed_kernel<true, T><<<g, 256, 0, s>>>((const T*) s0, (const T*) s1, (const T*) s2, c0, c1, c2, n, p);
		else mct3_fwd_packed_kernel<false, T><<<g, 256, 0, s>>>((const T*) s0, (const T*) s1, (const T*) s2, c0, c1, c2, n, p);
	});
}

void launch_mct_inv_packed(const int32_t *c0, const int32_t *c1, const int32_t *c2, void *d0, void *d1, void *d2, uint64_t n,
		const int32_t shift[3], const int32_t lo[3], const int32_t hi[3], int reversible, uint32_t sample_bytes, int sgnd, cudaStream_t s) {
	if (!n) return;
	Shift3 p;
	for (int i = 0; i < 3; ++i) { p.s[i] = shift[i]; p.lo[i] = lo[i]; p.hi[i] = hi[i]; }
	const unsigned g = grid_for(n / PK_UNROLL);
	with_sample_type(sample_bytes, sgnd, [&](auto t) {
		using T = decltype(t);
		if (reversible) mct3_inv_packed_kernel<true, T><<<g, 256, 0, s>>>(c0, c1, c2, (T*) d0, (T*) d1, (T*) d2, n, p);
		else mct3_inv_packed_kernel<false, T><<<g, 256, 0, s>>>(c0, c1, c2, (T*) d0, (T*) d1, (T*) d2, n, p);
	});
}

void launch_dcshift_fwd_packed(const void *src, int32_t *dst, uint64_t n, int32_t shift, int reversible, uint32_t sample_bytes, int sgnd,
		cudaStream_t s) {
	if (!n) return;
	const unsigned g = grid_for(n / PK_UNROLL);
	with_sample_type(sample_bytes, sgnd, [&](auto t) {
		using T = decltype(t);
		if (reversible) dcshift_packed_kernel<true, true, T><<<g, 256, 0, s>>>(src, dst, n, shift, 0, 0);
		else dcshift_packed_kernel<true, false, T><<<g, 256, 0, s>>>(src, dst, n, shift, 0, 0);
	});
}

void launch_dcshift_inv_packed(const int32_t *src, void *dst, uint64_t n, int32_t shift, int reversible, int32_t lo, int32_t hi,
		uint32_t sample_bytes, int sgnd, cudaStream_t s) {
	if (!n) return;
	const unsigned g = grid_for(n / PK_UNROLL);
	with_sample_type(sample_bytes, sgnd, [&](auto t) {
		using T = decltype(t);
		if (reversible) dcshift_packed_kernel<false, true, T><<<g, 256, 0, s>>>(src, dst, n, shift, lo, hi);
		else dcshift_packed_kernel<false, false, T><<<g, 256, 0, s>>>(src, dst, n, shift, lo, hi);
	});
}

} // namespace gb
