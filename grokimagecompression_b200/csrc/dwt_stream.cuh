// K2 / K3, second generation: streaming wavelet kernels, one WARP per work item, no barrier.
//
// forward 5/3  WaveletForward.h:40-161 + dwt53.cpp:150-169   (int32, exact)
// forward 9/7  WaveletForward.h:40-161 + dwt97.cpp:90-123    (int32 13-bit fixed point, exact)
// inverse 5/3  dwt.cpp:724-858, 256-363, 661-718             (int32, exact)
// inverse 9/7  dwt.cpp:1544-1738, 1413-1537, constants 172-178 (fp32, multiply-then-add, no FMA)
//
// A work item is a strip of the level's region: 128 columns (4 per lane; the outer lanes are the halo,
// 120 valid columns) by R rows (R chosen per launch by the plan, csrc/api.cu).  The warp walks down the strip two rows
// per trip:
//   forward: every lane loads its four columns of a low-pass and a high-pass row (one 16-byte load each,
//            the warp reads 512 contiguous bytes per row), advances the vertical lifting recurrences that it
//            keeps in registers (3 values per column for 9/7, 1 for 5/3, plus the previous row pair, which simply stays in
//            the prefetch queue), and for the two rows that
//            become final runs the horizontal lifting ACROSS LANES: a lane holds (low, high, low, high), the
//            neighbour's value arrives by one shuffle per lifting step, two rows per shuffle round.  The four
//            sub-band rows leave as 8-byte stores, 240 contiguous bytes per warp and sub-band.
//   inverse: the mirror image: sub-band rows in (8-byte loads), horizontal synthesis across lanes, vertical
//            synthesis recurrences in registers, 16-byte stores of finished rows.
// The vertical recurrences are the same arithmetic as lifting a window with a halo of 4 (9/7) / 2 (5/3)
// rows: a result row depends on source rows at distance <= halo, so a warp starts 2*LAG trips early and the
// first rows it emits are already exact.  Borders use whole-sample symmetric reflection of the source
// index, which reproduces the reference's clamped-neighbour code (see the note in dwt.cu); lines of length
// 1 are not lifted (dwt53.cpp:160, dwt.cpp:344-349, 1482-1490).  Rows are fetched G trips ahead, into registers
// (default) or into a per-warp ring in shared memory filled by cp.async; with no barrier every warp of the SM
// overlaps its loads with the others' arithmetic.  The level launches of a transform are chained by programmatic
// dependent launch, so the start-up of level l + 1 overlaps the tail of level l.
//
// This header is also compiled for the CPU by tests/dwt_emu.cpp (GB_EMU: one OS thread per lane), which checks
// the kernels against the oracle without a GPU.
#pragma once
#ifndef GB_EMU
#include "common.cuh"
#else
#include "dwt_plane.h"
#endif

namespace gb {

// halo lanes each side: 1 (120 valid columns per work item) or 2 (112 valid columns: every sub-band row segment a warp
// stores then starts on a 32-byte sector when the band does)
__host__ __device__ constexpr int dws_tw(int hl) { return 128 - 8 * hl; }
#ifndef DWS_WARPS_PER_CTA
#define DWS_WARPS_PER_CTA 4
#endif
#ifndef DWS_MINB
#define DWS_MINB 1   // __launch_bounds__ minimum CTAs per SM (register cap), a tuning knob
#endif
constexpr int DWS_WARPS = DWS_WARPS_PER_CTA; // work items per CTA

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

// asynchronous global -> shared copies (LDGSTS): the prefetch queue of a warp.  A lane only ever reads back the 16 bytes it
// copied itself, so cp.async.wait_group is all the synchronisation the queue needs.
#ifdef GB_EMU
__device__ __forceinline__ void dws_cp16(void *sm, const void *g) { DWS_ALIGNED(g, 16); memcpy(sm, g, 16); }
__device__ __forceinline__ void dws_cp8(void *sm, const void *g) { DWS_ALIGNED(g, 8); memcpy(sm, g, 8); }
__device__ __forceinline__ void dws_cp4(void *sm, const void *g) { memcpy(sm, g, 4); }
__device__ __forceinline__ void dws_commit() {}
template<int N> __device__ __forceinline__ void dws_wait() {}
#else
__device__ __forceinline__ void dws_cp16(void *sm, const void *g) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t) __cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void dws_cp8(void *sm, const void *g) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t) __cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void dws_cp4(void *sm, const void *g) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t) __cvta_generic_to_shared(sm)), "l"(g) : "memory");
}
__device__ __forceinline__ void dws_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void dws_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#endif

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
// (pa, pb): the row pair of the previous trip, a: this trip's low-pass row (its high-pass row is used one trip later)
template<> struct VFwd<true> {
	static constexpr int LAG = 1;
	int32_t pd;
	__device__ __forceinline__ void init() { pd = 0; }
	__device__ __forceinline__ void feed(int32_t pa, int32_t pb, int32_t a, int32_t &lo, int32_t &hi) {
		const int32_t dn = pb - ((pa + a) >> 1);
		lo = pa + ((pd + dn + 2) >> 2);
		hi = dn;
		pd = dn;
	}
};
template<> struct VFwd<false> {
	static constexpr int LAG = 2;
	int32_t pd1, ps1, pd2;
	__device__ __forceinline__ void init() { pd1 = ps1 = pd2 = 0; }
	__device__ __forceinline__ void feed(int32_t pa, int32_t pb, int32_t a, int32_t &lo, int32_t &hi) {
		const int32_t d1n = pb - dws_fix13(pa + a, 12994);
		const int32_t s1n = pa - dws_fix13(pd1 + d1n, 434);
		const int32_t d2n = pd1 + dws_fix13(ps1 + s1n, 7233);
		const int32_t s2n = ps1 + dws_fix13(pd2 + d2n, 3633);
		lo = dws_fix13(s2n, 6659);
		hi = dws_fix13(d2n, 5039);
		pd1 = d1n; ps1 = s1n; pd2 = d2n;
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

// Programmatic dependent launch: the level launches of a transform form a chain; every kernel lets its successor start
// at once (launch_dependents first thing), and the successor reads its tables, which no kernel writes, before it waits
// for the predecessor's results (griddepcontrol.wait = the whole predecessor grid has finished and its writes are visible).
// So launch latency and the start-up table reads of level l + 1 overlap the tail of level l.
#ifdef GB_EMU
__device__ __forceinline__ void dws_launch_dependents() {}
__device__ __forceinline__ void dws_grid_wait() {}
#else
__device__ __forceinline__ void dws_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void dws_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// decode the work item of this warp; returns false if the warp has nothing to do
struct DwsItem {
	int lane, X0, Y0, c; // c: region column of the lane's first element (a low-pass column)
};

__device__ __forceinline__ bool dws_item(const DwtPlane *__restrict__ planes, const uint32_t *__restrict__ item_plane, uint32_t nitems,
		int R, int hl, DwtPlane &P, DwsItem &it) {
	dws_launch_dependents();
	it.lane = threadIdx.x & 31;
	uint32_t item = blockIdx.x * DWS_WARPS + (threadIdx.x >> 5);
	if (item >= nitems) return false;
	P = planes[__ldg(item_plane + item)];
	item -= P.first_cta;
	it.X0 = (int) (item % P.tiles_x) * dws_tw(hl);
	it.Y0 = (int) (item / P.tiles_x) * R;
	it.c = it.X0 - 4 * hl - (int) P.cas_x + 4 * it.lane;
	dws_grid_wait(); // from here on the planes are read
	return true;
}

// =========================================================================================================
// forward
// =========================================================================================================
// two rows in lock step: twice the independent work between two shuffles
template<bool REV>
__device__ __forceinline__ void dws_hfwd2(Quad &p, Quad &q) {
	if (REV) {
		int32_t pe2 = dws_down(p.e0), qe2 = dws_down(q.e0);
		p.o0 -= (p.e0 + p.e1) >> 1; q.o0 -= (q.e0 + q.e1) >> 1;
		p.o1 -= (p.e1 + pe2) >> 1; q.o1 -= (q.e1 + qe2) >> 1;
		int32_t pom = dws_up(p.o1), qom = dws_up(q.o1);
		p.e0 += (pom + p.o0 + 2) >> 2; q.e0 += (qom + q.o0 + 2) >> 2;
		p.e1 += (p.o0 + p.o1 + 2) >> 2; q.e1 += (q.o0 + q.o1 + 2) >> 2;
	} else {
		int32_t pe2 = dws_down(p.e0), qe2 = dws_down(q.e0);
		p.o0 -= dws_fix13(p.e0 + p.e1, 12994); q.o0 -= dws_fix13(q.e0 + q.e1, 12994);
		p.o1 -= dws_fix13(p.e1 + pe2, 12994); q.o1 -= dws_fix13(q.e1 + qe2, 12994);
		int32_t pom = dws_up(p.o1), qom = dws_up(q.o1);
		p.e0 -= dws_fix13(pom + p.o0, 434); q.e0 -= dws_fix13(qom + q.o0, 434);
		p.e1 -= dws_fix13(p.o0 + p.o1, 434); q.e1 -= dws_fix13(q.o0 + q.o1, 434);
		pe2 = dws_down(p.e0); qe2 = dws_down(q.e0);
		p.o0 += dws_fix13(p.e0 + p.e1, 7233); q.o0 += dws_fix13(q.e0 + q.e1, 7233);
		p.o1 += dws_fix13(p.e1 + pe2, 7233); q.o1 += dws_fix13(q.e1 + qe2, 7233);
		pom = dws_up(p.o1); qom = dws_up(q.o1);
		p.e0 += dws_fix13(pom + p.o0, 3633); q.e0 += dws_fix13(qom + q.o0, 3633);
		p.e1 += dws_fix13(p.o0 + p.o1, 3633); q.e1 += dws_fix13(q.o0 + q.o1, 3633);
		p.e0 = dws_fix13(p.e0, 6659); p.e1 = dws_fix13(p.e1, 6659); q.e0 = dws_fix13(q.e0, 6659); q.e1 = dws_fix13(q.e1, 6659);
		p.o0 = dws_fix13(p.o0, 5039); p.o1 = dws_fix13(p.o1, 5039); q.o0 = dws_fix13(q.o0, 5039); q.o1 = dws_fix13(q.o1, 5039);
	}
}

__device__ __forceinline__ const int32_t *dws_at(const int32_t *p, uint32_t row, uint32_t stride_bytes) {
	return reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(p) + (size_t) row * stride_bytes); // one IMAD.WIDE.U32
}
__device__ __forceinline__ int32_t *dws_at(int32_t *p, uint32_t row, uint32_t stride_bytes) {
	return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(p) + (size_t) row * stride_bytes);
}

// One strip.  EDGE = false: every lane loads 16 aligned bytes of the region and every valid lane stores full aligned
// pairs (strips in the interior of a plane whose rows are 16-byte aligned): no predicates, no fallbacks in the loop.
// EDGE = true: the strip touches the left or right border of the region (or the plane is not aligned): every lane
// gathers its four (reflected) columns with 4-byte copies and stores element by element under predicates.
// Rows are fetched G trips (2 G rows) ahead of their use, into one of two prefetch queues:
// RING = false: row registers of the lane, filled by ordinary loads (nothing between DRAM and the registers); the default
//               for all four kernels: measured faster than the ring on every workload (profiles/README.md);
// RING = true : this warp's ring of 2 G rows of 32 x 16 bytes in shared memory, filled by asynchronous copies (no
//               registers held while the rows are in flight; GB200_DWT_RING=1).
template<bool REV, int G, bool EDGE, bool RING>
__device__ __forceinline__ void dws_fwd_strip(const DwtPlane &P, const DwsItem &it, int R, int hl, int4 *ring) {
	const int rw = (int) P.rw, rh = (int) P.rh, casx = (int) P.cas_x, casy = (int) P.cas_y;
	const int c = it.c;
	const uint32_t sstr4 = P.src_stride * 4u, dstr4 = P.dst_stride * 4u;

	const int32_t *s0 = P.src + (EDGE ? dws_reflect(c, rw) : c);
	const int32_t *s1 = P.src + dws_reflect(c + 1, rw), *s2 = P.src + dws_reflect(c + 2, rw), *s3 = P.src + dws_reflect(c + 3, rw);
	int4 *const slot0 = ring + it.lane;
	// RING = false: the queue.  It has one row pair more than the G that are in flight: the vertical pass needs the previous
	// trip's rows once more, so they stay where they were loaded and their slot is refilled one trip later (no copies; the slot
	// numbers are compile-time constants once the trip loop is unrolled).
	constexpr int K = RING ? G : G + 1; // row pairs in the queue; trip t lives in pair t % K
	Quad regs[RING ? 1 : 2 * K];
	auto issue_row = [&](int y, int slot) {
		const uint32_t gy = (uint32_t) dws_reflect(y, rh);
		if (RING) {
			int4 *sm = slot0 + 32 * slot;
			if (!EDGE) dws_cp16(sm, dws_at(s0, gy, sstr4));
			else {
				int32_t *e = reinterpret_cast<int32_t*>(sm);
				dws_cp4(e, dws_at(s0, gy, sstr4)); dws_cp4(e + 1, dws_at(s1, gy, sstr4));
				dws_cp4(e + 2, dws_at(s2, gy, sstr4)); dws_cp4(e + 3, dws_at(s3, gy, sstr4));
			}
		} else {
			Quad &q = regs[RING ? 0 : slot];
			if (!EDGE) {
				const int4 v = dws_ld4(dws_at(s0, gy, sstr4));
				q.e0 = v.x; q.o0 = v.y; q.e1 = v.z; q.o1 = v.w;
			} else {
				q.e0 = dws_ld1(dws_at(s0, gy, sstr4)); q.o0 = dws_ld1(dws_at(s1, gy, sstr4));
				q.e1 = dws_ld1(dws_at(s2, gy, sstr4)); q.o1 = dws_ld1(dws_at(s3, gy, sstr4));
			}
		}
	};
	auto read_row = [&](int slot) -> Quad {
		if (!RING) return regs[RING ? 0 : slot];
		const int4 v = slot0[32 * slot];
		Quad q;
		q.e0 = v.x; q.o0 = v.y; q.e1 = v.z; q.o1 = v.w;
		return q;
	};

	// destination columns: the element at region column cc goes to index cc >> 1 of its sub-band
	const bool lv = it.lane >= hl && it.lane <= 31 - hl;
	const bool okE0 = lv && (unsigned) c < (unsigned) rw, okO0 = lv && (unsigned) (c + 1) < (unsigned) rw;
	const bool okE1 = lv && (unsigned) (c + 2) < (unsigned) rw, okO1 = lv && (unsigned) (c + 3) < (unsigned) rw;
	int32_t *const dL = P.dst + (c >> 1), *const dH = P.dst + ((int) P.sw + ((c + 1) >> 1)); // second element of each kind: + 1
	auto store_row = [&](const Quad &q, uint32_t row) {
		int32_t *oL = dws_at(dL, row, dstr4), *oH = dws_at(dH, row, dstr4);
		if (!EDGE) {
			if (lv) { dws_st2(oL, q.e0, q.e1); dws_st2(oH, q.o0, q.o1); }
		} else {
			if (okE0) dws_st1(oL, q.e0);
			if (okE1) dws_st1(oL + 1, q.e1);
			if (okO0) dws_st1(oH, q.o0);
			if (okO1) dws_st1(oH + 1, q.o1);
		}
	};
	const bool hlift = !EDGE || rw > 1;
	auto emit1 = [&](Quad q, int y, bool high_row) {
		if (hlift) dws_hfwd<REV>(q);
		else if (REV && casx) { q.e0 *= 2; q.o0 *= 2; q.e1 *= 2; q.o1 *= 2; } // dwt53.cpp:160
		store_row(q, (uint32_t) ((y >> 1) + (high_row ? (int) P.sh : 0)));
	};

	if (rh == 1) { // a single row is not lifted vertically; with an odd origin it is a high-pass row (x2 for 5/3)
		if (it.Y0 != 0) return;
		issue_row(0, 0);
		if (RING) { dws_commit(); dws_wait<0>(); }
		Quad q = read_row(0);
		if (REV && casy) { q.e0 *= 2; q.o0 *= 2; q.e1 *= 2; q.o1 *= 2; }
		emit1(q, 0, casy != 0);
		return;
	}

	constexpr int LAG = VFwd<REV>::LAG;
	const int ys = it.Y0 - casy - 2 * LAG;              // first row fed (a low-pass row)
	const int yv0 = max(it.Y0 - casy, 0), yv1 = min(it.Y0 - casy + R, rh); // rows this item owns
	if (yv1 <= yv0) return;
	const int niter = LAG + ((yv1 - 1 - ys) >> 1) + 1;  // trip j emits rows ys + 2 (j - LAG) and the next one

	VFwd<REV> v0, v1, v2, v3;
	v0.init(); v1.init(); v2.init(); v3.init();
	#pragma unroll
	for (int g = 0; g < G; ++g) { // one copy group per trip (two rows), G trips in flight
		if (g < niter) { issue_row(ys + 2 * g, 2 * g); issue_row(ys + 2 * g + 1, 2 * g + 1); }
		if (RING) dws_commit();
	}
	Quad pa = {0, 0, 0, 0}, pb = {0, 0, 0, 0}; // RING: the previous trip's rows (copies)
	if (!RING) { regs[RING ? 0 : 2 * (K - 1)] = pa; regs[RING ? 0 : 2 * (K - 1) + 1] = pb; } // "trip -1": part of the warm-up
	for (int j0 = 0; j0 < niter; j0 += K) {
		#pragma unroll
		for (int u = 0; u < K; ++u) {
			const int j = j0 + u;
			if (j >= niter) break;
			const int prev = (u + K - 1) % K; // pair of trip j - 1
			if (RING) dws_wait<G - 1>();
			const Quad a = read_row(2 * u), b = read_row(2 * u + 1);
			const Quad qa = RING ? pa : read_row(2 * prev), qb = RING ? pb : read_row(2 * prev + 1);
			Quad lo, hi;
			v0.feed(qa.e0, qb.e0, a.e0, lo.e0, hi.e0);
			v1.feed(qa.o0, qb.o0, a.o0, lo.o0, hi.o0);
			v2.feed(qa.e1, qb.e1, a.e1, lo.e1, hi.e1);
			v3.feed(qa.o1, qb.o1, a.o1, lo.o1, hi.o1);
			// refill: RING overwrites the pair just read (it was copied out), the register queue the pair of trip j - 1
			if (j + G < niter) { issue_row(ys + 2 * (j + G), RING ? 2 * u : 2 * prev); issue_row(ys + 2 * (j + G) + 1, RING ? 2 * u + 1 : 2 * prev + 1); }
			if (RING) { dws_commit(); pa = a; pb = b; }
			// rows before the first owned one come out of the warm-up trips and fail the range test, as do rows past the last
			const int yl = ys + 2 * (j - LAG);
			const bool vl = yl >= yv0 && yl < yv1, vh = yl + 1 >= yv0 && yl + 1 < yv1;
			if (vl && vh && hlift) {
				dws_hfwd2<REV>(lo, hi);
				store_row(lo, (uint32_t) (yl >> 1));
				store_row(hi, (uint32_t) (((yl + 1) >> 1) + (int) P.sh));
			} else {
				if (vl) emit1(lo, yl, false);
				if (vh) emit1(hi, yl + 1, true);
			}
		}
	}
}

template<bool REV, int G, bool RING>
__global__ void __launch_bounds__(DWS_WARPS * 32, DWS_MINB) dwt_fwd_stream_kernel(const DwtPlane *__restrict__ planes,
		const uint32_t *__restrict__ item_plane, uint32_t nitems, int R, int hl) {
	__shared__ int4 ring[RING ? DWS_WARPS : 1][RING ? 2 * G : 1][RING ? 32 : 1];
	DwtPlane P;
	DwsItem it;
	if (!dws_item(planes, item_plane, nitems, R, hl, P, it)) return;
	const int c = it.c, rw = (int) P.rw;
	const bool lv = it.lane >= hl && it.lane <= 31 - hl;
	const bool vld = c >= 0 && c + 3 < rw && (c & 3) == 0 && (P.src_stride & 3) == 0 && (((size_t) P.src) & 15) == 0;
	const bool vst = (P.dst_stride & 1) == 0 && (((size_t) P.dst) & 7) == 0 && ((c >> 1) & 1) == 0 && ((P.sw + ((c + 1) >> 1)) & 1) == 0;
	int4 *const my_ring = &ring[RING ? threadIdx.x >> 5 : 0][0][0];
	if (__all_sync(0xffffffffu, vld && (vst || !lv))) dws_fwd_strip<REV, G, false, RING>(P, it, R, hl, my_ring);
	else dws_fwd_strip<REV, G, true, RING>(P, it, R, hl, my_ring);
}

// =========================================================================================================
// inverse
// =========================================================================================================
template<bool REV>
__device__ __forceinline__ void dws_hinv2(Quad &p, Quad &q) {
	if (REV) {
		int32_t pom = dws_up(p.o1), qom = dws_up(q.o1);
		p.e0 -= (pom + p.o0 + 2) >> 2; q.e0 -= (qom + q.o0 + 2) >> 2;
		p.e1 -= (p.o0 + p.o1 + 2) >> 2; q.e1 -= (q.o0 + q.o1 + 2) >> 2;
		int32_t pe2 = dws_down(p.e0), qe2 = dws_down(q.e0);
		p.o0 += (p.e0 + p.e1) >> 1; q.o0 += (q.e0 + q.e1) >> 1;
		p.o1 += (p.e1 + pe2) >> 1; q.o1 += (q.e1 + qe2) >> 1;
	} else {
		float pe0 = __fmul_rn(dws_f(p.e0), DWS_KL), pe1 = __fmul_rn(dws_f(p.e1), DWS_KL), qe0 = __fmul_rn(dws_f(q.e0), DWS_KL), qe1 = __fmul_rn(dws_f(q.e1), DWS_KL);
		float po0 = __fmul_rn(dws_f(p.o0), DWS_KH), po1 = __fmul_rn(dws_f(p.o1), DWS_KH), qo0 = __fmul_rn(dws_f(q.o0), DWS_KH), qo1 = __fmul_rn(dws_f(q.o1), DWS_KH);
		float pom = dws_f(dws_up(dws_i(po1))), qom = dws_f(dws_up(dws_i(qo1)));
		pe0 = dws_step(pe0, pom, po0, DWS_C1); qe0 = dws_step(qe0, qom, qo0, DWS_C1);
		pe1 = dws_step(pe1, po0, po1, DWS_C1); qe1 = dws_step(qe1, qo0, qo1, DWS_C1);
		float pe2 = dws_f(dws_down(dws_i(pe0))), qe2 = dws_f(dws_down(dws_i(qe0)));
		po0 = dws_step(po0, pe0, pe1, DWS_C2); qo0 = dws_step(qo0, qe0, qe1, DWS_C2);
		po1 = dws_step(po1, pe1, pe2, DWS_C2); qo1 = dws_step(qo1, qe1, qe2, DWS_C2);
		pom = dws_f(dws_up(dws_i(po1))); qom = dws_f(dws_up(dws_i(qo1)));
		pe0 = dws_step(pe0, pom, po0, DWS_C3); qe0 = dws_step(qe0, qom, qo0, DWS_C3);
		pe1 = dws_step(pe1, po0, po1, DWS_C3); qe1 = dws_step(qe1, qo0, qo1, DWS_C3);
		pe2 = dws_f(dws_down(dws_i(pe0))); qe2 = dws_f(dws_down(dws_i(qe0)));
		po0 = dws_step(po0, pe0, pe1, DWS_C4); qo0 = dws_step(qo0, qe0, qe1, DWS_C4);
		po1 = dws_step(po1, pe1, pe2, DWS_C4); qo1 = dws_step(qo1, qe1, qe2, DWS_C4);
		p.e0 = dws_i(pe0); p.o0 = dws_i(po0); p.e1 = dws_i(pe1); p.o1 = dws_i(po1);
		q.e0 = dws_i(qe0); q.o0 = dws_i(qo0); q.e1 = dws_i(qe1); q.o1 = dws_i(qo1);
	}
}

// EDGE as in the forward kernel: false = every lane loads aligned pairs of all four sub-bands and every valid lane stores
// 16 aligned bytes; true = 4-byte gathers through reflected indices, element-wise predicated stores.  Queue slot of a lane:
// (low, low, high, high) = the two 8-byte pairs as they lie in the sub-bands.
template<bool REV, int G, bool EDGE, bool RING>
__device__ __forceinline__ void dws_inv_strip(const DwtPlane &P, const DwsItem &it, int R, int hl, int4 *ring) {
	const int rw = (int) P.rw, rh = (int) P.rh, casx = (int) P.cas_x, casy = (int) P.cas_y;
	const int c = it.c;
	const uint32_t sstr4 = P.src_stride * 4u, bstr4 = P.band_stride * 4u, dstr4 = P.dst_stride * 4u;

	// coefficient columns: the sample at region column g is index g >> 1 of the low-pass (LL / LH) or high-pass (HL / HH) band
	const int k0 = (EDGE ? dws_reflect(c, rw) : c) >> 1, k1 = (int) P.sw + ((EDGE ? dws_reflect(c + 1, rw) : c + 1) >> 1);
	const int k2 = dws_reflect(c + 2, rw) >> 1, k3 = (int) P.sw + (dws_reflect(c + 3, rw) >> 1);
	const int32_t *const pLL = P.src + k0, *const pLL2 = P.src + k2;   // vertical low-pass rows: LL (previous level's output) | HL
	const int32_t *const pLH = P.band + k0, *const pLH2 = P.band + k2; // vertical high-pass rows: LH | HH
	const int32_t *const pH = P.band + k1, *const pH2 = P.band + k3;   // HL and HH columns
	int4 *const slot0 = ring + it.lane;
	Quad regs[RING ? 1 : 2 * G];
	auto issue_row = [&](int y, bool high, int slot) {
		const uint32_t k = (uint32_t) (dws_reflect(y, rh) >> 1) + (high ? P.sh : 0u);
		const int32_t *pl = high ? pLH : pLL, *pl2 = high ? pLH2 : pLL2;
		const uint32_t lstr4 = high ? bstr4 : sstr4;
		if (RING) {
			int32_t *e = reinterpret_cast<int32_t*>(slot0 + 32 * slot);
			if (!EDGE) { dws_cp8(e, dws_at(pl, k, lstr4)); dws_cp8(e + 2, dws_at(pH, k, bstr4)); }
			else {
				dws_cp4(e, dws_at(pl, k, lstr4)); dws_cp4(e + 1, dws_at(pl2, k, lstr4));
				dws_cp4(e + 2, dws_at(pH, k, bstr4)); dws_cp4(e + 3, dws_at(pH2, k, bstr4));
			}
		} else {
			Quad &q = regs[RING ? 0 : slot];
			if (!EDGE) {
				const int2 v = dws_ld2(dws_at(pl, k, lstr4)), w = dws_ld2(dws_at(pH, k, bstr4));
				q.e0 = v.x; q.e1 = v.y; q.o0 = w.x; q.o1 = w.y;
			} else {
				q.e0 = dws_ld1(dws_at(pl, k, lstr4)); q.e1 = dws_ld1(dws_at(pl2, k, lstr4));
				q.o0 = dws_ld1(dws_at(pH, k, bstr4)); q.o1 = dws_ld1(dws_at(pH2, k, bstr4));
			}
		}
	};
	auto read_row = [&](int slot) -> Quad {
		if (!RING) return regs[RING ? 0 : slot];
		const int4 v = slot0[32 * slot];
		Quad q;
		q.e0 = v.x; q.e1 = v.y; q.o0 = v.z; q.o1 = v.w;
		return q;
	};
	const bool hlift = !EDGE || rw > 1;
	auto hsyn1 = [&](Quad &q) {
		if (hlift) dws_hinv<REV>(q);
		else if (REV && casx) { q.e0 /= 2; q.o0 /= 2; q.e1 /= 2; q.o1 /= 2; } // dwt.cpp:344-349 (C division)
	};

	const bool lv = it.lane >= hl && it.lane <= 31 - hl;
	const bool ok0 = lv && (unsigned) c < (unsigned) rw, ok1 = lv && (unsigned) (c + 1) < (unsigned) rw;
	const bool ok2 = lv && (unsigned) (c + 2) < (unsigned) rw, ok3 = lv && (unsigned) (c + 3) < (unsigned) rw;
	int32_t *const d0 = P.dst + c;
	auto store_row = [&](const Quad &q, int y) {
		int32_t *o = dws_at(d0, (uint32_t) y, dstr4);
		if (!EDGE) {
			if (lv) dws_st4(o, q.e0, q.o0, q.e1, q.o1);
		} else {
			if (ok0) dws_st1(o, q.e0);
			if (ok1) dws_st1(o + 1, q.o0);
			if (ok2) dws_st1(o + 2, q.e1);
			if (ok3) dws_st1(o + 3, q.o1);
		}
	};

	if (rh == 1) { // a single row: low-pass (LL | HL) with an even origin, high-pass with an odd one; no vertical step
		if (it.Y0 != 0) return;
		issue_row(0, casy != 0, 0);
		if (RING) { dws_commit(); dws_wait<0>(); }
		Quad q = read_row(0);
		hsyn1(q);
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
	// rows ys + 2 j are vertical low-pass rows, ys + 2 j + 1 high-pass ones (reflection keeps the parity)
	#pragma unroll
	for (int g = 0; g < G; ++g) {
		if (g < niter) { issue_row(ys + 2 * g, false, 2 * g); issue_row(ys + 2 * g + 1, true, 2 * g + 1); }
		if (RING) dws_commit();
	}
	for (int j0 = 0; j0 < niter; j0 += G) {
		#pragma unroll
		for (int u = 0; u < G; ++u) {
			const int j = j0 + u;
			if (j >= niter) break;
			if (RING) dws_wait<G - 1>();
			Quad a = read_row(2 * u), b = read_row(2 * u + 1);
			if (RING) { // the ring slot was copied out by the read: refill it at once
				if (j + G < niter) { issue_row(ys + 2 * (j + G), false, 2 * u); issue_row(ys + 2 * (j + G) + 1, true, 2 * u + 1); }
				dws_commit();
			}
			if (hlift) dws_hinv2<REV>(a, b);
			else { hsyn1(a); hsyn1(b); }
			// register queue: the raw rows are dead once the horizontal synthesis has consumed them; reloading their registers
			// only now saves the copies a refill before that would force
			if (!RING && j + G < niter) { issue_row(ys + 2 * (j + G), false, 2 * u); issue_row(ys + 2 * (j + G) + 1, true, 2 * u + 1); }
			Quad r0, r1;
			v0.feed(a.e0, b.e0, r0.e0, r1.e0);
			v1.feed(a.o0, b.o0, r0.o0, r1.o0);
			v2.feed(a.e1, b.e1, r0.e1, r1.e1);
			v3.feed(a.o1, b.o1, r0.o1, r1.o1);
			const int yl = ys + 2 * (j - LAG);
			if (yl >= yv0 && yl < yv1) store_row(r0, yl);
			if (yl + 1 >= yv0 && yl + 1 < yv1) store_row(r1, yl + 1);
		}
	}
}

template<bool REV, int G, bool RING>
__global__ void __launch_bounds__(DWS_WARPS * 32, DWS_MINB) dwt_inv_stream_kernel(const DwtPlane *__restrict__ planes,
		const uint32_t *__restrict__ item_plane, uint32_t nitems, int R, int hl) {
	__shared__ int4 ring[RING ? DWS_WARPS : 1][RING ? 2 * G : 1][RING ? 32 : 1];
	DwtPlane P;
	DwsItem it;
	if (!dws_item(planes, item_plane, nitems, R, hl, P, it)) return;
	const int c = it.c, rw = (int) P.rw;
	const bool inside = c >= 0 && c + 3 < rw && (c & 3) == 0; // then the low-pass index c / 2 is even
	const bool vld = inside && (P.sw & 1) == 0 && ((P.src_stride | P.band_stride) & 1) == 0 && ((((size_t) P.src) | ((size_t) P.band)) & 7) == 0;
	const bool vst = (P.dst_stride & 3) == 0 && (((size_t) P.dst) & 15) == 0;
	int4 *const my_ring = &ring[RING ? threadIdx.x >> 5 : 0][0][0];
	if (__all_sync(0xffffffffu, vld && vst)) dws_inv_strip<REV, G, false, RING>(P, it, R, hl, my_ring);
	else dws_inv_strip<REV, G, true, RING>(P, it, R, hl, my_ring);
}

} // namespace gb
