// Test infrastructure (CPU only): runs the streaming DWT kernels of csrc/dwt_stream.cuh on the host, one OS
// thread per lane (shuffles = exchange through a 32-party barrier), with the same level tables the plan
// builds in csrc/api.cu, and compares every result with the oracle (oracle/libgb_oracle.so).  It lets the
// kernels' index arithmetic (parities, reflections, vector / scalar paths, strips and row chunks) be checked
// where there is no GPU; the GPU parity tests in tests/test_gpu_stages.py remain the proof for the device.
//
//   g++ -std=c++20 -O1 -ffp-contract=off -pthread -DGB_EMU tests/dwt_emu.cpp -Loracle -lgb_oracle -o tests/_dwt_emu
#include <barrier>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include <algorithm>

// ---- the handful of CUDA names the kernels use -----------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
struct Idx { unsigned x = 0, y = 0, z = 0; };
static thread_local Idx threadIdx, blockIdx;
static thread_local int emu_lane;
static std::barrier<> emu_bar(32);
static int emu_buf[32];
static inline int __shfl_down_sync(unsigned, int v, int d) {
	emu_buf[emu_lane] = v;
	emu_bar.arrive_and_wait();
	int r = emu_lane + d < 32 ? emu_buf[emu_lane + d] : v;
	emu_bar.arrive_and_wait();
	return r;
}
static inline int __shfl_up_sync(unsigned, int v, int d) {
	emu_buf[emu_lane] = v;
	emu_bar.arrive_and_wait();
	int r = emu_lane - d >= 0 ? emu_buf[emu_lane - d] : v;
	emu_bar.arrive_and_wait();
	return r;
}
static inline int __all_sync(unsigned, int pred) {
	emu_buf[emu_lane] = pred;
	emu_bar.arrive_and_wait();
	int r = 1;
	for (int i = 0; i < 32; ++i) r &= emu_buf[i] != 0;
	emu_bar.arrive_and_wait();
	return r;
}
template<typename T> static inline T __ldg(const T *p) { return *p; }
template<typename T> static inline void __stcg(T *p, T v) { *p = v; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
using std::max;
using std::min;
static const int32_t *emu_lo, *emu_hi; // every access must stay inside the planes
static inline void gb_emu_check_aligned(const void *p, int n) {
	if (((uintptr_t) p) % n) { fprintf(stderr, "misaligned %d-byte access\n", n); abort(); }
}

#include "../grokimagecompression_b200/csrc/dwt_stream.cuh"

extern "C" int gbo_dwt_fwd(int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres, int reversible);
extern "C" int gbo_dwt_inv(int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres, uint32_t numres_decode,
		int reversible);

using gb::DwtPlane;

template<typename K>
static void launch(K kernel, const DwtPlane *planes, const uint32_t *map, uint32_t nitems, int R, int hl) {
	const unsigned grid = (nitems + gb::DWS_WARPS - 1) / gb::DWS_WARPS;
	std::vector<std::thread> th;
	for (int lane = 0; lane < 32; ++lane)
		th.emplace_back([=] {
			emu_lane = lane;
			for (unsigned b = 0; b < grid; ++b)
				for (unsigned w = 0; w < (unsigned) gb::DWS_WARPS; ++w) {
					blockIdx.x = b;
					threadIdx.x = w * 32 + lane;
					kernel(planes, map, nitems, R, hl);
				}
		});
	for (auto &t : th) t.join();
}

static uint32_t cdiv2n(uint32_t a, uint32_t n) { return (uint32_t) (((uint64_t) a + (1ull << n) - 1) >> n); }

// the plan's table for one level of one plane (csrc/api.cu)
static int g_hl = 1, g_ring = 1;
static void one_level(DwtPlane d, bool fwd, int rev, int R) {
	const int hl = g_hl;
	d.tiles_x = (d.rw + d.cas_x + gb::dws_tw(hl) - 1) / gb::dws_tw(hl);
	d.tiles_y = (d.rh + d.cas_y + R - 1) / R;
	d.first_cta = 0;
	if (!d.rw || !d.rh) return;
	const uint32_t n = d.tiles_x * d.tiles_y;
	std::vector<uint32_t> map(n, 0);
	if (fwd) {
		if (rev) { if (g_ring) launch(gb::dwt_fwd_stream_kernel<true, 4, true>, &d, map.data(), n, R, hl); else launch(gb::dwt_fwd_stream_kernel<true, 2, false>, &d, map.data(), n, R, hl); }
		else { if (g_ring) launch(gb::dwt_fwd_stream_kernel<false, 8, true>, &d, map.data(), n, R, hl); else launch(gb::dwt_fwd_stream_kernel<false, 3, false>, &d, map.data(), n, R, hl); }
	} else {
		if (rev) { if (g_ring) launch(gb::dwt_inv_stream_kernel<true, 8, true>, &d, map.data(), n, R, hl); else launch(gb::dwt_inv_stream_kernel<true, 4, false>, &d, map.data(), n, R, hl); }
		else { if (g_ring) launch(gb::dwt_inv_stream_kernel<false, 4, true>, &d, map.data(), n, R, hl); else launch(gb::dwt_inv_stream_kernel<false, 2, false>, &d, map.data(), n, R, hl); }
	}
}

struct Geom { uint32_t x0, y0, x1, y1, nr; };

// planes are allocated 16-byte aligned with some slack, like the plan's (256-byte aligned, back to back)
static int32_t *plane_alloc(size_t n) {
	void *p = aligned_alloc(256, ((n + 64) * 4 + 255) / 256 * 256);
	memset(p, 0, ((n + 64) * 4 + 255) / 256 * 256);
	return (int32_t*) p;
}

static int check_fwd(const Geom &g, int rev, int R, std::mt19937 &rng) {
	const uint32_t w = g.x1 - g.x0, h = g.y1 - g.y0;
	std::vector<int32_t> ref((size_t) w * h), out((size_t) w * h);
	std::uniform_int_distribution<int> dist(-32768, 32767);
	for (auto &v : ref) v = dist(rng) * (rev ? 1 : 8);
	int32_t *pl[2] = {plane_alloc((size_t) w * h), plane_alloc((size_t) w * h)};
	memcpy(pl[0], ref.data(), ref.size() * 4);
	out = ref;
	const uint32_t levels = g.nr - 1;
	for (uint32_t i = 0; i < levels; ++i) {
		DwtPlane d;
		memset(&d, 0, sizeof(d));
		d.rw = cdiv2n(g.x1, i) - cdiv2n(g.x0, i); d.rh = cdiv2n(g.y1, i) - cdiv2n(g.y0, i);
		d.sw = cdiv2n(g.x1, i + 1) - cdiv2n(g.x0, i + 1); d.sh = cdiv2n(g.y1, i + 1) - cdiv2n(g.y0, i + 1);
		d.cas_x = cdiv2n(g.x0, i) & 1; d.cas_y = cdiv2n(g.y0, i) & 1;
		d.src_stride = d.band_stride = d.dst_stride = w;
		d.src = pl[i & 1]; d.dst = pl[(i & 1) ^ 1];
		one_level(d, true, rev, R);
		if (d.rw && d.rh)
			for (uint32_t y = 0; y < d.rh; ++y) memcpy(&out[(size_t) y * w], d.dst + (size_t) y * w, d.rw * 4);
	}
	gbo_dwt_fwd(ref.data(), g.x0, g.y0, g.x1, g.y1, g.nr, rev);
	free(pl[0]); free(pl[1]);
	size_t bad = 0, first = 0;
	for (size_t i = 0; i < ref.size(); ++i) if (ref[i] != out[i]) { if (!bad) first = i; ++bad; }
	if (bad) printf("FWD MISMATCH rev=%d R=%d geom=(%u,%u,%u,%u,%u): %zu samples, first at x=%zu y=%zu\n", rev, R, g.x0, g.y0, g.x1, g.y1, g.nr,
			bad, first % w, first / w);
	return bad != 0;
}

static int check_inv(const Geom &g, uint32_t nd, int rev, int R, std::mt19937 &rng) {
	const uint32_t top = g.nr - nd;
	const uint32_t w = cdiv2n(g.x1, top) - cdiv2n(g.x0, top), h = cdiv2n(g.y1, top) - cdiv2n(g.y0, top);
	std::vector<int32_t> ref((size_t) w * h), out;
	std::uniform_int_distribution<int> dist(-4096, 4095);
	std::normal_distribution<float> nd_(0.f, 100.f);
	for (auto &v : ref) { if (rev) v = dist(rng); else { float f = nd_(rng); memcpy(&v, &f, 4); } }
	int32_t *pl[3] = {plane_alloc((size_t) w * h), plane_alloc((size_t) w * h), plane_alloc((size_t) w * h)};
	memcpy(pl[0], ref.data(), ref.size() * 4);
	int final_role = 0;
	const uint32_t levels = nd - 1;
	for (uint32_t i = 0; i < levels; ++i) {
		const uint32_t lvl = g.nr - 2 - i;
		DwtPlane d;
		memset(&d, 0, sizeof(d));
		d.rw = cdiv2n(g.x1, lvl) - cdiv2n(g.x0, lvl); d.rh = cdiv2n(g.y1, lvl) - cdiv2n(g.y0, lvl);
		d.sw = cdiv2n(g.x1, lvl + 1) - cdiv2n(g.x0, lvl + 1); d.sh = cdiv2n(g.y1, lvl + 1) - cdiv2n(g.y0, lvl + 1);
		d.cas_x = cdiv2n(g.x0, lvl) & 1; d.cas_y = cdiv2n(g.y0, lvl) & 1;
		d.src_stride = d.band_stride = d.dst_stride = w;
		const int src_role = i == 0 ? 0 : (i & 1 ? 1 : 2), dst_role = i & 1 ? 2 : 1;
		d.src = pl[src_role]; d.band = pl[0]; d.dst = pl[dst_role];
		final_role = dst_role;
		one_level(d, false, rev, R);
	}
	out.assign(pl[final_role], pl[final_role] + (size_t) w * h);
	gbo_dwt_inv(ref.data(), g.x0, g.y0, g.x1, g.y1, g.nr, nd, rev);
	for (auto p : pl) free(p);
	size_t bad = 0, first = 0;
	for (size_t i = 0; i < ref.size(); ++i) if (ref[i] != out[i]) { if (!bad) first = i; ++bad; }
	if (bad) printf("INV MISMATCH rev=%d R=%d geom=(%u,%u,%u,%u,%u) nd=%u: %zu samples, first at x=%zu y=%zu\n", rev, R, g.x0, g.y0, g.x1,
			g.y1, g.nr, nd, bad, first % w, first / w);
	return bad != 0;
}

int main(int argc, char **argv) {
	const bool quick = argc > 1 && !strcmp(argv[1], "quick");
	if (argc > 1 && !strcmp(argv[1], "fuzz")) { // dwt_emu fuzz <cases> <seed>: random origins, sizes, levels, rows per item, queues
		const int n = argc > 2 ? atoi(argv[2]) : 100;
		std::mt19937 rng(argc > 3 ? atoi(argv[3]) : 1);
		auto U = [&](int lo, int hi) { return (uint32_t) std::uniform_int_distribution<int>(lo, hi)(rng); };
		int fails = 0;
		for (int i = 0; i < n; ++i) {
			Geom g;
			g.x0 = U(0, 1) ? U(0, 9) : 0; g.y0 = U(0, 1) ? U(0, 9) : 0;
			g.x1 = g.x0 + (U(0, 3) ? U(1, 300) : U(1, 9)); g.y1 = g.y0 + (U(0, 3) ? U(1, 150) : U(1, 9));
			if (U(0, 4) == 0) g.x1 = g.x0 + 4 * U(30, 130); // aligned widths reach the interior (vector) strips
			g.nr = U(1, 6);
			const int rev = (int) U(0, 1), R = (int) (2 * U(1, 40)); // the plan picks 2 rows per item for the smallest levels
			g_hl = (int) U(1, 2); g_ring = (int) U(0, 1);
			fails += check_fwd(g, rev, R, rng);
			fails += check_inv(g, U(1, g.nr), rev, R, rng);
		}
		printf("fuzz: %d cases, %d failed\n", 2 * n, fails);
		return fails ? 1 : 0;
	}
	// the geometries of tests/test_gpu_stages.py plus a few that cross strip / chunk borders
	std::vector<Geom> geoms = {{0, 0, 64, 64, 6}, {0, 0, 37, 53, 4}, {3, 5, 40, 41, 6}, {1, 1, 2, 2, 3}, {7, 0, 8, 33, 5}, {0, 0, 1, 1, 2},
		{5, 3, 300, 211, 6}, {1, 0, 3, 1, 3}, {0, 0, 129, 65, 3}, {63, 63, 64 + 130, 64 + 67, 4}, {0, 0, 5, 1, 4}, {0, 0, 1, 7, 4},
		{9, 9, 10, 200, 3}, {0, 0, 120, 16, 2}, {1, 1, 121, 17, 2}, {1, 1, 122, 18, 3}, {0, 0, 241, 40, 3}, {0, 0, 2, 2, 2}, {0, 1, 3, 4, 3},
		{0, 0, 256, 48, 4}, {2, 2, 250, 35, 4}};
	if (!quick) { geoms.push_back({0, 0, 1024, 112, 6}); geoms.push_back({1024, 2048, 2048, 2160, 6}); }
	std::mt19937 rng(5);
	int fails = 0, cases = 0;
	for (int rev = 1; rev >= 0; --rev)
		for (const Geom &g : geoms)
			for (int R : {2, 4, 10, 16, 64}) {
				if (quick && (R == 10 || R == 4) && (uint64_t) (g.x1 - g.x0) * (g.y1 - g.y0) > 6000) continue; // keep the CPU suite short
				if (quick && R == 2 && (uint64_t) (g.x1 - g.x0) * (g.y1 - g.y0) > 20000) continue;
				g_hl = (R == 16 || R == 2) ? 1 : 2;
				g_ring = R != 16 && R != 2;
				fails += check_fwd(g, rev, R, rng); ++cases;
				std::vector<uint32_t> nds = {g.nr, std::max(1u, g.nr - 1), g.nr > 3 ? g.nr - 3 : 1u, 1u};
				std::sort(nds.begin(), nds.end());
				nds.erase(std::unique(nds.begin(), nds.end()), nds.end());
				for (uint32_t nd : nds) { fails += check_inv(g, nd, rev, R, rng); ++cases; }
			}
	printf("%d cases, %d failed\n", cases, fails);
	return fails ? 1 : 0;
}
