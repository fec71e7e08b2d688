// K2 / K3, second generation: streaming wavelet kernels, one WARP per work item, no shared memory.
//
// forward 5/3  WaveletForward.h:40-161 + dwt53.cpp:150-169   (int32, exact)
// forward 9/7  WaveletForward.h:40-161 + dwt97.cpp:90-123    (int32 13-bit fixed point, exact)
// inverse 5/3  dwt.cpp:724-858, 256-363, 661-718             (int32, exact)
// inverse 9/7  dwt.cpp:1544-1738, 1413-1537, constants 172-178 (fp32, multiply-then-add, no FMA)
//
// A work item is a strip of the level's region: 128 columns (4 per lane; the outer lanes are the halo,
// 120 valid columns) by R rows.  The warp walks down the strip two rows per trip:
//   forward: every lane loads its four columns of a low-pass and a high-pass row (one 16-byte load each,
//            the warp reads 512 contiguous bytes per row), advances the vertical lifting recurrences that it
//            keeps in registers (5 values per column for 9/7, 3 for 5/3), and for the two rows that
//            become final runs the horizontal lifting ACROSS LANES: a lane holds (low, high, low, high), the
//            neighbour's value arrives by one shuffle per lifting step.  The four sub-band rows leave as
//            8-byte stores, 240 contiguous bytes per warp and sub-band.
//   inverse: the mirror image: sub-band rows in (8-byte loads), horizontal synthesis across lanes, vertical
//            synthesis recurrences in registers, 16-byte stores of finished rows.
// The vertical recurrences are the same arithmetic as lifting a window with a halo of 4 (9/7) / 2 (5/3)
// rows: a result row depends on source rows at distance <= halo, so a warp starts 2*LAG trips early and the
// first rows it emits are already exact.  Borders use whole-sample symmetric reflection of the source
// index, which reproduces the reference's clamped-neighbour code (see the note in dwt.cu); lines of length
// 1 are not lifted (dwt53.cpp:160, dwt.cpp:344-349, 1482-1490).  Rows are prefetched U trips ahead in
// registers; with no barrier and no shared memory every warp of the SM overlaps its loads with the others'
// arithmetic.
//
// This header is also compiled for the CPU by tests/dwt_emu (GB_EMU: one OS thread per lane), which checks
// the kernels against the oracle without a GPU.
#pragma once
#ifndef GB_EMU
#include "common.cuh"
#else
#include "dwt_plane.h"
#endif

namespace gb {

constexpr int DWS_TW = 120;      // valid columns per work item (lanes 1..30)
constexpr int DWS_WARPS = 4;     // work items per CTA

__device__ __forceinline__ int dws_reflect(int i, int len) {
	if ((unsigned) i < (unsigned) len) return i;
	if (len == 1) return 0;
	const int p = 2 * (len - 1);
	i %= p;
	if (i < 0) i += p;
	return i >= len ? p - i : i;
}

__device__ __forceinline__ int32_t dws_fix13(int32_t a, int32_t b) {
	return (int32_t) (((int64_t) a * (int64_t) b + 4096) >> 13);
}

struct Quad { int32_t e0, o0, e1, o1; }; // four consecutive columns of a row: low, high, low, high

#ifdef GB_EMU
#define DWS_ALIGNED(p, n) gb_emu_check_aligned((const void*) (p), n)
#else
#define DWS_ALIGNED(p, n) ((void) 0)
#endif
// launches are out of place, so every load may take the read-only path; results are stored to L2 only (the next
// level / Tier-1 reads them from there).  The explicit intrinsics also tell the compiler these are global addresses
// (the pointers come out of a table, it would emit generic loads and stores otherwise).
__device__ __forceinline__ int4 dws_ld4(const int32_t *p) { DWS_ALIGNED(p, 16); return __ldg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ int2 dws_ld2(const int32_t *p) { DWS_ALIGNED(p, 8); return __ldg(reinterpret_cast<const int2*>(p)); }
__device__ __forceinline__ int32_t dws_ld1(const int32_t *p) { return __ldg(p); }
__device__ __forceinline__ void dws_st4(int32_t *p, int32_t a, int32_t b, int32_t c, int32_t d) { DWS_ALIGNED(p, 16); __stcg(reinterpret_cast<int4*>(p), make_int4(a, b, c, d)); }
__device__ __forceinline__ void dws_st2(int32_t *p, int32_t a, int32_t b) { DWS_ALIGNED(p, 8); __stcg(reinterpret_cast<int2*>(p), make_int2(a, b)); }
__device__ __forceinline__ void dws_st1(int32_t *p, int32_t a) { __stcg(p, a); }

__device__ __forceinline__ int32_t dws_down(int32_t v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ int32_t dws_up(int32_t v) { return __shfl_up_sync(0xffffffffu, v, 1); }

// ---- horizontal lifting of one row spread over the warp; lanes 1..30 come out exact ------------------
template<bool REV>
__device__ __forceinline__ void dws_hfwd(Quad &q) {
	if (REV) {
		int32_t e2 = dws_down(q.e0);
		q.o0 -= (q.e0 + q.e1) >> 1;
		q.o1 -= (q.e1 + e2) >> 1;
		int32_t om = dws_up(q.o1);
		q.e0 += (om + q.o0 + 2) >> 2;
		q.e1 += (q.o0 + q.o1 + 2) >> 2;
	} else {
		int32_t e2 = dws_down(q.e0);
		q.o0 -= dws_fix13(q.e0 + q.e1, 12994);
		q.o1 -= dws_fix13(q.e1 + e2, 12994);
		int32_t om = dws_up(q.o1);
		q.e0 -= dws_fix13(om + q.o0, 434);
		q.e1 -= dws_fix13(q.o0 + q.o1, 434);
		e2 = dws_down(q.e0);
		q.o0 += dws_fix13(q.e0 + q.e1, 7233);
		q.o1 += dws_fix13(q.e1 + e2, 7233);
		om = dws_up(q.o1);
		q.e0 += dws_fix13(om + q.o0, 3633);
		q.e1 += dws_fix13(q.o0 + q.o1, 3633);
		q.e0 = dws_fix13(q.e0, 6659); q.e1 = dws_fix13(q.e1, 6659);
		q.o0 = dws_fix13(q.o0, 5039); q.o1 = dws_fix13(q.o1, 5039);
	}
}

__device__ __forceinline__ float dws_f(int32_t v) { return __int_as_float(v); }
__device__ __forceinline__ int32_t dws_i(float v) { return __float_as_int(v); }
// x + (a + b) * c, multiply then add (dwt.cpp:1413-1471)
__device__ __forceinline__ float dws_step(float x, float a, float b, float c) { return __fadd_rn(x, __fmul_rn(__fadd_rn(a, b), c)); }

constexpr float DWS_KL = 1.230174105f, DWS_KH = 1.625732422f; // dwt.cpp:172-178
constexpr float DWS_C1 = -0.443506852f, DWS_C2 = -0.882911075f, DWS_C3 = 0.052980118f, DWS_C4 = 1.586134342f;

template<bool REV>
__device__ __forceinline__ void dws_hinv(Quad &q) {
	if (REV) {
		int32_t om = dws_up(q.o1);
		q.e0 -= (om + q.o0 + 2) >> 2;
		q.e1 -= (q.o0 + q.o1 + 2) >> 2;
		int32_t e2 = dws_down(q.e0);
		q.o0 += (q.e0 + q.e1) >> 1;
		q.o1 += (q.e1 + e2) >> 1;
	} else {
		float e0 = __fmul_rn(dws_f(q.e0), DWS_KL), e1 = __fmul_rn(dws_f(q.e1), DWS_KL);
		float o0 = __fmul_rn(dws_f(q.o0), DWS_KH), o1 = __fmul_rn(dws_f(q.o1), DWS_KH);
		float om = dws_f(dws_up(dws_i(o1)));
		e0 = dws_step(e0, om, o0, DWS_C1); e1 = dws_step(e1, o0, o1, DWS_C1);
		float e2 = dws_f(dws_down(dws_i(e0)));
		o0 = dws_step(o0, e0, e1, DWS_C2); o1 = dws_step(o1, e1, e2, DWS_C2);
		om = dws_f(dws_up(dws_i(o1)));
		e0 = dws_step(e0, om, o0, DWS_C3); e1 = dws_step(e1, o0, o1, DWS_C3);
		e2 = dws_f(dws_down(dws_i(e0)));
		o0 = dws_step(o0, e0, e1, DWS_C4); o1 = dws_step(o1, e1, e2, DWS_C4);
		q.e0 = dws_i(e0); q.o0 = dws_i(o0); q.e1 = dws_i(e1); q.o1 = dws_i(o1);
	}
}

// ---- vertical recurrences, one column ------------------------------------------------------------------
// forward: feed (a, b) = next low-pass and high-pass source rows; returns the low / high result rows that
// became final: stream index j - LAG when fed pair j.
template<bool REV> struct VFwd;
template<> struct VFwd<true> {
	static constexpr int LAG = 1;
	int32_t pa, pb, pd;
	__device__ __forceinline__ void init() { pa = pb = pd = 0; }
	__device__ __forceinline__ void feed(int32_t a, int32_t b, int32_t &lo, int32_t &hi) {
		const int32_t dn = pb - ((pa + a) >> 1);
		lo = pa + ((pd + dn + 2) >> 2);
		hi = dn;
		pa = a; pb = b; pd = dn;
	}
};
template<> struct VFwd<false> {
	static constexpr int LAG = 2;
	int32_t pa, pb, pd1, ps1, pd2;
	__device__ __forceinline__ void init() { pa = pb = pd1 = ps1 = pd2 = 0; }
	__device__ __forceinline__ void feed(int32_t a, int32_t b, int32_t &lo, int32_t &hi) {
		const int32_t d1n = pb - dws_fix13(pa + a, 12994);
		const int32_t s1n = pa - dws_fix13(pd1 + d1n, 434);
		const int32_t d2n = pd1 + dws_fix13(ps1 + s1n, 7233);
		const int32_t s2n = ps1 + dws_fix13(pd2 + d2n, 3633);
		lo = dws_fix13(s2n, 6659);
		hi = dws_fix13(d2n, 5039);
		pa = a; pb = b; pd1 = d1n; ps1 = s1n; pd2 = d2n;
	}
};

// inverse: feed (l, h) = next low-pass and high-pass coefficient rows (after the horizontal synthesis);
// returns the two finished image rows of stream index j - LAG.
template<bool REV> struct VInv;
template<> struct VInv<true> {
	static constexpr int LAG = 1;
	int32_t ph, ps;
	__device__ __forceinline__ void init() { ph = ps = 0; }
	__device__ __forceinline__ void feed(int32_t l, int32_t h, int32_t &r0, int32_t &r1) {
		const int32_t sn = l - ((ph + h + 2) >> 2);
		r0 = ps;
		r1 = ph + ((ps + sn) >> 1);
		ph = h; ps = sn;
	}
};
template<> struct VInv<false> {
	static constexpr int LAG = 2;
	float ph, ps1, pd1, ps2;
	__device__ __forceinline__ void init() { ph = ps1 = pd1 = ps2 = 0.f; }
	__device__ __forceinline__ void feed(int32_t li, int32_t hi, int32_t &r0, int32_t &r1) {
		const float l = __fmul_rn(dws_f(li), DWS_KL), h = __fmul_rn(dws_f(hi), DWS_KH);
		const float s1n = dws_step(l, ph, h, DWS_C1);
		const float d1n = dws_step(ph, ps1, s1n, DWS_C2);
		const float s2n = dws_step(ps1, pd1, d1n, DWS_C3);
		const float d2n = dws_step(pd1, ps2, s2n, DWS_C4);
		r0 = dws_i(ps2);
		r1 = dws_i(d2n);
		ph = h; ps1 = s1n; pd1 = d1n; ps2 = s2n;
	}
};

// decode the work item of this warp; returns false if the warp has nothing to do
struct DwsItem {
	int lane, X0, Y0, c; // c: region column of the lane's first element (a low-pass column)
};

__device__ __forceinline__ bool dws_item(const DwtPlane *__restrict__ planes, const uint32_t *__restrict__ item_plane, uint32_t nitems,
		int R, DwtPlane &P, DwsItem &it) {
	it.lane = threadIdx.x & 31;
	uint32_t item = blockIdx.x * DWS_WARPS + (threadIdx.x >> 5);
	if (item >= nitems) return false;
	P = planes[__ldg(item_plane + item)];
	item -= P.first_cta;
	it.X0 = (int) (item % P.tiles_x) * DWS_TW;
	it.Y0 = (int) (item / P.tiles_x) * R;
	it.c = it.X0 - 4 - (int) P.cas_x + 4 * it.lane;
	return true;
}

// =========================================================================================================
// forward
// =========================================================================================================
template<bool REV, int U>
__global__ void __launch_bounds__(DWS_WARPS * 32) dwt_fwd_stream_kernel(const DwtPlane *__restrict__ planes,
		const uint32_t *__restrict__ item_plane, uint32_t nitems, int R) {
	DwtPlane P;
	DwsItem it;
	if (!dws_item(planes, item_plane, nitems, R, P, it)) return;
	const int rw = (int) P.rw, rh = (int) P.rh, casx = (int) P.cas_x, casy = (int) P.cas_y;
	const int c = it.c;
	const size_t sstr = P.src_stride, dstr = P.dst_stride;

	// source columns (reflected) and the 16-byte fast path
	const int g0 = dws_reflect(c, rw), g1 = dws_reflect(c + 1, rw), g2 = dws_reflect(c + 2, rw), g3 = dws_reflect(c + 3, rw);
	const bool vld = c >= 0 && c + 3 < rw && (c & 3) == 0 && (P.src_stride & 3) == 0 && (((size_t) P.src) & 15) == 0;
	auto load_row = [&](int y) -> Quad {
		const int32_t *p = P.src + (size_t) dws_reflect(y, rh) * sstr;
		Quad q;
		if (vld) {
			const int4 v = dws_ld4(p + c);
			q.e0 = v.x; q.o0 = v.y; q.e1 = v.z; q.o1 = v.w;
		} else {
			q.e0 = dws_ld1(p + g0); q.o0 = dws_ld1(p + g1); q.e1 = dws_ld1(p + g2); q.o1 = dws_ld1(p + g3);
		}
		return q;
	};

	// destination columns: element at region column cc goes to index cc >> 1 of its sub-band
	const bool lv = it.lane >= 1 && it.lane <= 30;
	const bool okE0 = lv && (unsigned) c < (unsigned) rw, okO0 = lv && (unsigned) (c + 1) < (unsigned) rw;
	const bool okE1 = lv && (unsigned) (c + 2) < (unsigned) rw, okO1 = lv && (unsigned) (c + 3) < (unsigned) rw;
	const int iL = c >> 1, iH = (int) P.sw + ((c + 1) >> 1); // second element of each kind: + 1
	const bool al = (P.dst_stride & 1) == 0 && (((size_t) P.dst) & 7) == 0;
	const bool vL = okE0 && okE1 && al && (iL & 1) == 0, vH = okO0 && okO1 && al && (iH & 1) == 0;

	auto emit = [&](Quad q, int y, bool high_row) {
		if (rw > 1) dws_hfwd<REV>(q);
		else if (REV && casx) { q.e0 *= 2; q.o0 *= 2; q.e1 *= 2; q.o1 *= 2; } // dwt53.cpp:160
		int32_t *o = P.dst + (size_t) ((y >> 1) + (high_row ? (int) P.sh : 0)) * dstr;
		if (vL) dws_st2(o + iL, q.e0, q.e1);
		else { if (okE0) dws_st1(o + iL, q.e0); if (okE1) dws_st1(o + iL + 1, q.e1); }
		if (vH) dws_st2(o + iH, q.o0, q.o1);
		else { if (okO0) dws_st1(o + iH, q.o0); if (okO1) dws_st1(o + iH + 1, q.o1); }
	};

	if (rh == 1) { // a single row is not lifted vertically; with an odd origin it is a high-pass row (x2 for 5/3)
		Quad q = load_row(0);
		if (REV && casy) { q.e0 *= 2; q.o0 *= 2; q.e1 *= 2; q.o1 *= 2; }
		emit(q, 0, casy != 0);
		return;
	}

	constexpr int LAG = VFwd<REV>::LAG;
	const int ys = it.Y0 - casy - 2 * LAG;              // first row fed (a low-pass row)
	const int yv0 = max(it.Y0 - casy, 0), yv1 = min(it.Y0 - casy + R, rh); // rows this item owns
	if (yv1 <= yv0) return;
	const int niter = LAG + ((yv1 - 1 - ys) >> 1) + 1;  // trip j emits rows ys + 2 (j - LAG) and the next one

	VFwd<REV> v0, v1, v2, v3;
	v0.init(); v1.init(); v2.init(); v3.init();
	Quad nxt[2 * U];
	#pragma unroll
	for (int u = 0; u < 2 * U; ++u) nxt[u] = load_row(ys + u);
	for (int j0 = 0; j0 < niter; j0 += U) {
		Quad cur[2 * U];
		#pragma unroll
		for (int u = 0; u < 2 * U; ++u) cur[u] = nxt[u];
		if (j0 + U < niter) {
			#pragma unroll
			for (int u = 0; u < 2 * U; ++u) nxt[u] = load_row(ys + 2 * (j0 + U) + u);
		}
		#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int j = j0 + u;
			if (j < niter) {
				const Quad a = cur[2 * u], b = cur[2 * u + 1];
				Quad lo, hi;
				v0.feed(a.e0, b.e0, lo.e0, hi.e0);
				v1.feed(a.o0, b.o0, lo.o0, hi.o0);
				v2.feed(a.e1, b.e1, lo.e1, hi.e1);
				v3.feed(a.o1, b.o1, lo.o1, hi.o1);
				const int yl = ys + 2 * (j - LAG);
				if (j >= 2 * LAG) {
					if (yl >= yv0 && yl < yv1) emit(lo, yl, false);
					if (yl + 1 >= yv0 && yl + 1 < yv1) emit(hi, yl + 1, true);
				}
			}
		}
	}
}

// =========================================================================================================
// inverse
// =========================================================================================================
template<bool REV, int U>
__global__ void __launch_bounds__(DWS_WARPS * 32) dwt_inv_stream_kernel(const DwtPlane *__restrict__ planes,
		const uint32_t *__restrict__ item_plane, uint32_t nitems, int R) {
	DwtPlane P;
	DwsItem it;
	if (!dws_item(planes, item_plane, nitems, R, P, it)) return;
	const int rw = (int) P.rw, rh = (int) P.rh, casx = (int) P.cas_x, casy = (int) P.cas_y;
	const int c = it.c;
	const size_t sstr = P.src_stride, bstr = P.band_stride, dstr = P.dst_stride;

	// coefficient columns: the sample at region column g is index g >> 1 of the low-pass (LL / LH) or high-pass (HL / HH) band
	const int k0 = dws_reflect(c, rw) >> 1, k1 = (int) P.sw + (dws_reflect(c + 1, rw) >> 1);
	const int k2 = dws_reflect(c + 2, rw) >> 1, k3 = (int) P.sw + (dws_reflect(c + 3, rw) >> 1);
	const bool inside = c >= 0 && c + 3 < rw;
	const bool vsrc = inside && (k0 & 1) == 0 && (P.src_stride & 1) == 0 && (((size_t) P.src) & 7) == 0;
	const bool bal = (P.band_stride & 1) == 0 && (((size_t) P.band) & 7) == 0;
	const bool vbl = inside && (k0 & 1) == 0 && bal, vbh = inside && (k1 & 1) == 0 && bal;
	auto load_row = [&](int y) -> Quad {
		const int gy = dws_reflect(y, rh);
		const int k = gy >> 1;
		const bool high = ((gy + casy) & 1) != 0;
		Quad q;
		if (!high) { // LL from the previous level's output, HL from the coefficient plane
			const int32_t *pl = P.src + (size_t) k * sstr, *ph = P.band + (size_t) k * bstr;
			if (vsrc) { const int2 v = dws_ld2(pl + k0); q.e0 = v.x; q.e1 = v.y; }
			else { q.e0 = dws_ld1(pl + k0); q.e1 = dws_ld1(pl + k2); }
			if (vbh) { const int2 v = dws_ld2(ph + k1); q.o0 = v.x; q.o1 = v.y; }
			else { q.o0 = dws_ld1(ph + k1); q.o1 = dws_ld1(ph + k3); }
		} else { // LH and HH
			const int32_t *pb = P.band + (size_t) ((int) P.sh + k) * bstr;
			if (vbl) { const int2 v = dws_ld2(pb + k0); q.e0 = v.x; q.e1 = v.y; }
			else { q.e0 = dws_ld1(pb + k0); q.e1 = dws_ld1(pb + k2); }
			if (vbh) { const int2 v = dws_ld2(pb + k1); q.o0 = v.x; q.o1 = v.y; }
			else { q.o0 = dws_ld1(pb + k1); q.o1 = dws_ld1(pb + k3); }
		}
		return q;
	};
	// horizontal synthesis of a loaded row (at consumption time, so that the prefetch does not wait for its loads)
	auto hsyn = [&](Quad &q) {
		if (rw > 1) dws_hinv<REV>(q);
		else if (REV && casx) { q.e0 /= 2; q.o0 /= 2; q.e1 /= 2; q.o1 /= 2; } // dwt.cpp:344-349 (C division)
	};

	const bool lv = it.lane >= 1 && it.lane <= 30;
	const bool ok0 = lv && (unsigned) c < (unsigned) rw, ok1 = lv && (unsigned) (c + 1) < (unsigned) rw;
	const bool ok2 = lv && (unsigned) (c + 2) < (unsigned) rw, ok3 = lv && (unsigned) (c + 3) < (unsigned) rw;
	const bool vst = ok0 && ok3 && (c & 3) == 0 && (P.dst_stride & 3) == 0 && (((size_t) P.dst) & 15) == 0;
	auto store_row = [&](const Quad &q, int y) {
		int32_t *o = P.dst + (size_t) y * dstr + c;
		if (vst) dws_st4(o, q.e0, q.o0, q.e1, q.o1);
		else { if (ok0) dws_st1(o, q.e0); if (ok1) dws_st1(o + 1, q.o0); if (ok2) dws_st1(o + 2, q.e1); if (ok3) dws_st1(o + 3, q.o1); }
	};

	if (rh == 1) {
		Quad q = load_row(0);
		hsyn(q);
		if (REV && casy) { q.e0 /= 2; q.o0 /= 2; q.e1 /= 2; q.o1 /= 2; }
		store_row(q, 0);
		return;
	}

	constexpr int LAG = VInv<REV>::LAG;
	const int ys = it.Y0 - casy - 2 * LAG;
	const int yv0 = max(it.Y0 - casy, 0), yv1 = min(it.Y0 - casy + R, rh);
	if (yv1 <= yv0) return;
	const int niter = LAG + ((yv1 - 1 - ys) >> 1) + 1;

	VInv<REV> v0, v1, v2, v3;
	v0.init(); v1.init(); v2.init(); v3.init();
	Quad nxt[2 * U];
	#pragma unroll
	for (int u = 0; u < 2 * U; ++u) nxt[u] = load_row(ys + u);
	for (int j0 = 0; j0 < niter; j0 += U) {
		Quad cur[2 * U];
		#pragma unroll
		for (int u = 0; u < 2 * U; ++u) cur[u] = nxt[u];
		if (j0 + U < niter) {
			#pragma unroll
			for (int u = 0; u < 2 * U; ++u) nxt[u] = load_row(ys + 2 * (j0 + U) + u);
		}
		#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int j = j0 + u;
			if (j < niter) {
				Quad a = cur[2 * u], b = cur[2 * u + 1];
				hsyn(a); hsyn(b);
				Quad r0, r1;
				v0.feed(a.e0, b.e0, r0.e0, r1.e0);
				v1.feed(a.o0, b.o0, r0.o0, r1.o0);
				v2.feed(a.e1, b.e1, r0.e1, r1.e1);
				v3.feed(a.o1, b.o1, r0.o1, r1.o1);
				const int yl = ys + 2 * (j - LAG);
				if (j >= 2 * LAG) {
					if (yl >= yv0 && yl < yv1) store_row(r0, yl);
					if (yl + 1 >= yv0 && yl + 1 < yv1) store_row(r1, yl + 1);
				}
			}
		}
	}
}

} // namespace gb
