// K2 / K3: forward and inverse wavelet transforms, one decomposition level per launch, every
// lifting step of both directions in one pass over the data.
//
// forward 5/3  WaveletForward.h:40-161 + dwt53.cpp:150-169   (int32, exact)
// forward 9/7  WaveletForward.h:40-161 + dwt97.cpp:90-123    (int32 13-bit fixed point, exact)
// inverse 5/3  dwt.cpp:724-858, 256-363, 661-718             (int32, exact)
// inverse 9/7  dwt.cpp:1544-1738, 1413-1537, constants 172-178 (fp32, multiply-then-add, no FMA)
//
// Two generations live here.  The product path is the STREAMING kernels of dwt_stream.cuh (one warp per strip, vertical
// lifting recurrences in registers, horizontal lifting across lanes, no shared memory, no barrier); this file holds
// their launch code.  The FIRST generation below (a CTA stages a 128-column tile in shared memory, thread = column,
// then thread = half a row) is kept behind GB200_DWT_LEGACY=1 as the A/B baseline of tools/dwt_bench.py:
// B200, configs[1] planes 9/7: 0.135 -> 0.116 ms forward, 0.367 -> 0.094 ms inverse; configs[2] planes 5/3: 0.467 -> 0.42 ms
// forward (78 % of the measured HBM copy peak), 1.90 -> 0.40 ms inverse (82 %).
//
// First generation: a CTA of 128 threads owns a tile of the interleaved (spatial) domain, 128 columns wide INCLUDING
// the halo of 2 (5/3) or 4 (9/7) samples each side, 64 rows high plus halo.  Lifting runs in REGISTERS, fully unrolled:
//   forward: thread = column; it loads its 64+2H samples straight from global memory (every load of
//            the warp is one coalesced row segment, all loads in flight at once), lifts vertically,
//            parks the 64 valid rows in shared memory; then thread = half a row lifts horizontally
//            from shared memory; the four sub-bands leave through shared memory as coalesced rows.
//   inverse: the four sub-bands are interleaved into shared memory (coalesced per sub-band row),
//            thread = half a row lifts horizontally, thread = column lifts vertically and stores its
//            64 rows coalesced.
// Launches are out of place (ping-pong planes): no CTA reads what another CTA of the launch writes.
//
// Borders use whole-sample symmetric reflection of the source index.  Lifting on a symmetrically
// extended signal keeps it symmetric, so this produces exactly the reference's clamped-neighbour
// results (the (c+c)*x special case of dwt.cpp:1463-1469 rounds like (x+x)*c).  Lines of length 1
// are the exception and are handled explicitly (dwt53.cpp:160, dwt.cpp:344-349, 1482-1490).
#include "common.cuh"
#include "dwt_stream.cuh"

namespace gb {

constexpr int NCOL = 128; // columns staged per CTA, halo included

// TH = valid rows per CTA: 64 for the big levels (least halo work), 32 / 16 for the small ones, where a
// launch is a single partial wave and the latency of one CTA is what the level costs
template<bool REV, int TH> struct Geo {
	static constexpr int H = REV ? 2 : 4;
	static constexpr int TW = NCOL - 2 * H;   // valid columns per CTA
	static constexpr int NR = TH + 2 * H;     // rows held per column
	static constexpr int HALF = TW / 2;       // valid columns per horizontal work item
	static constexpr int NU = HALF + 2 * H;   // samples held per horizontal work item
};

void dwt_tile_shape(int reversible, uint32_t *tw) {
	*tw = reversible ? Geo<true, 64>::TW : Geo<false, 64>::TW;
}

__device__ __forceinline__ int reflect(int i, int len) {
	// whole-sample symmetric extension, any distance
	if (len == 1) return 0;
	int p = 2 * (len - 1);
	i %= p;
	if (i < 0) i += p;
	return i >= len ? p - i : i;
}

// same, for an index at most len-1 outside [0,len) (one reflection is enough)
__device__ __forceinline__ int reflect1(int i, int len) {
	i = i < 0 ? -i : i;
	return i >= len ? 2 * (len - 1) - i : i;
}

__device__ __forceinline__ int32_t fix13(int32_t a, int32_t b) {
	return (int32_t) (((int64_t) a * (int64_t) b + 4096) >> 13);
}

// cta_plane[blockIdx.x] = index of the plane this CTA works on (built by the host with the launch table)
__device__ __forceinline__ DwtPlane find_plane(const DwtPlane *planes, const uint32_t *cta_plane, uint32_t &cta) {
	const DwtPlane P = planes[cta_plane[cta]];
	cta -= P.first_cta;
	return P;
}

// ---- lifting on a register array; HP = parity (index & 1) of the high-pass positions --------------
// After the call positions [H, N-H) are final; the outer H positions are scratch.

template<int N, int HP>
__device__ __forceinline__ void fwd53(int32_t (&x)[N]) {
	#pragma unroll
	for (int i = 1; i <= N - 2; ++i) if ((i & 1) == HP) x[i] -= (x[i - 1] + x[i + 1]) >> 1;
	#pragma unroll
	for (int i = 2; i <= N - 3; ++i) if ((i & 1) != HP) x[i] += (x[i - 1] + x[i + 1] + 2) >> 2;
}

template<int N, int HP>
__device__ __forceinline__ void inv53(int32_t (&x)[N]) {
	#pragma unroll
	for (int i = 1; i <= N - 2; ++i) if ((i & 1) != HP) x[i] -= (x[i - 1] + x[i + 1] + 2) >> 2;
	#pragma unroll
	for (int i = 2; i <= N - 3; ++i) if ((i & 1) == HP) x[i] += (x[i - 1] + x[i + 1]) >> 1;
}

template<int N, int HP>
__device__ __forceinline__ void fwd97(int32_t (&x)[N]) {
	#pragma unroll
	for (int i = 1; i <= N - 2; ++i) if ((i & 1) == HP) x[i] -= fix13(x[i - 1] + x[i + 1], 12994);
	#pragma unroll
	for (int i = 2; i <= N - 3; ++i) if ((i & 1) != HP) x[i] -= fix13(x[i - 1] + x[i + 1], 434);
	#pragma unroll
	for (int i = 3; i <= N - 4; ++i) if ((i & 1) == HP) x[i] += fix13(x[i - 1] + x[i + 1], 7233);
	#pragma unroll
	for (int i = 4; i <= N - 5; ++i) if ((i & 1) != HP) x[i] += fix13(x[i - 1] + x[i + 1], 3633);
	#pragma unroll
	for (int i = 4; i <= N - 5; ++i) x[i] = fix13(x[i], (i & 1) == HP ? 5039 : 6659);
}

template<int N, int HP>
__device__ __forceinline__ void inv97(int32_t (&xi)[N]) {
	float x[N];
	#pragma unroll
	for (int i = 0; i < N; ++i) x[i] = __fmul_rn(__int_as_float(xi[i]), (i & 1) == HP ? 1.625732422f : 1.230174105f);
	#pragma unroll
	for (int i = 1; i <= N - 2; ++i) if ((i & 1) != HP) x[i] = __fadd_rn(x[i], __fmul_rn(__fadd_rn(x[i - 1], x[i + 1]), -0.443506852f));
	#pragma unroll
	for (int i = 2; i <= N - 3; ++i) if ((i & 1) == HP) x[i] = __fadd_rn(x[i], __fmul_rn(__fadd_rn(x[i - 1], x[i + 1]), -0.882911075f));
	#pragma unroll
	for (int i = 3; i <= N - 4; ++i) if ((i & 1) != HP) x[i] = __fadd_rn(x[i], __fmul_rn(__fadd_rn(x[i - 1], x[i + 1]), 0.052980118f));
	#pragma unroll
	for (int i = 4; i <= N - 5; ++i) if ((i & 1) == HP) x[i] = __fadd_rn(x[i], __fmul_rn(__fadd_rn(x[i - 1], x[i + 1]), 1.586134342f));
	#pragma unroll
	for (int i = 0; i < N; ++i) xi[i] = __float_as_int(x[i]);
}

// len  : length of the lifted line in the region (1 = degenerate: no lifting)
// hp   : parity of the high-pass positions in the register array
// cas  : parity of the line's first sample on the canvas (1: a lone sample is high-pass)
template<bool FWD, bool REV, int N>
__device__ __forceinline__ void lift_line(int32_t (&x)[N], int len, int hp, int cas) {
	if (len > 1) {
		if (FWD) {
			if (REV) { if (hp) fwd53<N, 1>(x); else fwd53<N, 0>(x); }
			else { if (hp) fwd97<N, 1>(x); else fwd97<N, 0>(x); }
		} else {
			if (REV) { if (hp) inv53<N, 1>(x); else inv53<N, 0>(x); }
			else { if (hp) inv97<N, 1>(x); else inv97<N, 0>(x); }
		}
	} else if (REV && cas) {
		#pragma unroll
		for (int i = 0; i < N; ++i) x[i] = FWD ? x[i] * 2 : x[i] / 2; // dwt53.cpp:160 / dwt.cpp:349 (C division)
	}
}

template<bool REV, int TH>
__global__ void __launch_bounds__(NCOL) dwt_fwd_kernel(const DwtPlane *__restrict__ planes, const uint32_t *__restrict__ cta_plane) {
	using G = Geo<REV, TH>;
	constexpr int H = G::H, TW = G::TW, NR = G::NR, HALF = G::HALF, NU = G::NU, PITCH = NCOL + 1;
	__shared__ int32_t sm[TH * PITCH];

	uint32_t cta = blockIdx.x;
	const DwtPlane P = find_plane(planes, cta_plane, cta);
	const int X0 = (cta % P.tiles_x) * TW, Y0 = (cta / P.tiles_x) * TH;
	const int rw = P.rw, rh = P.rh;
	const int t = threadIdx.x;
	// local index l of a line <-> region index origin + l - H; origin and H are even, so the local
	// parity of the high-pass samples is 1 - cas
	const int hpx = 1 - (int) P.cas_x, hpy = 1 - (int) P.cas_y;

	{ // vertical (WaveletForward.h:91-121): thread = column
		const int gx = reflect(X0 - H + t, rw);
		const int32_t *col = P.src + gx;
		int32_t v[NR];
		if (Y0 - H >= 0 && Y0 - H + NR <= rh) { // interior rows: no reflection
			const int32_t *p = col + (size_t) (Y0 - H) * P.src_stride;
			#pragma unroll
			for (int r = 0; r < NR; ++r) v[r] = p[(size_t) r * P.src_stride];
		} else if (rh > 2 * H) { // border tile of a line longer than the halo: a single reflection
			#pragma unroll
			for (int r = 0; r < NR; ++r) v[r] = col[(size_t) reflect1(min(Y0 - H + r, rh + H - 1), rh) * P.src_stride];
		} else {
			#pragma unroll
			for (int r = 0; r < NR; ++r) v[r] = col[(size_t) reflect(Y0 - H + r, rh) * P.src_stride];
		}
		lift_line<true, REV, NR>(v, rh, hpy, (int) P.cas_y);
		#pragma unroll
		for (int r = 0; r < TH; ++r) sm[r * PITCH + t] = v[r + H];
	}
	__syncthreads();
	{ // horizontal (WaveletForward.h:124-152): work item = half a row; results go back de-interleaved
		const bool active = t < 2 * TH;
		const int row = active ? t >> 1 : 0, c0 = (t & 1) * HALF;
		int32_t u[NU];
		#pragma unroll
		for (int i = 0; i < NU; ++i) u[i] = sm[row * PITCH + c0 + i];
		lift_line<true, REV, NU>(u, rw, hpx, (int) P.cas_x);
		__syncthreads();
		// valid samples i in [H, H+HALF): tile column c0 + i - H; low-pass first, then high-pass
		if (active) {
			#pragma unroll
			for (int i = H; i < H + HALF; ++i) {
				const int lc = c0 + i - H;
				const int dst = (lc >> 1) + (((i & 1) == hpx) ? TW / 2 : 0);
				sm[row * PITCH + dst] = u[i];
			}
		}
	}
	__syncthreads();
	// de-interleave rows and store (dwt_utils.cpp:84-127): shared row r, column c holds the sample of
	// tile row r and, horizontally, low-pass index c (c < TW/2) or high-pass index c - TW/2
	const int lowx = (X0 >> 1), lowy = (Y0 >> 1);
	// number of valid low/high columns of this tile
	const int vw = min(TW, rw - X0), vh = min(TH, rh - Y0);
	const int nlow_x = (vw + (hpx ? 1 : 0)) >> 1, nhigh_x = vw - nlow_x;
	// warp w stores tile rows w, w+4, ...; a lane owns low-pass columns lane, lane+32 and the same two
	// high-pass columns: four predicated stores per row, predicates and offsets fixed per lane
	const int lane = t & 31, warp = t >> 5;
	const bool l0 = lane < nlow_x, l1 = lane + 32 < nlow_x, h0 = lane < nhigh_x, h1 = lane + 32 < nhigh_x;
	const int hoff = (int) P.sw + lane;
	#pragma unroll 4
	for (int r = warp; r < vh; r += NCOL / 32) {
		const bool hy = (r & 1) == hpy;
		int32_t *orow = P.dst + (size_t) (lowy + (r >> 1) + (hy ? (int) P.sh : 0)) * P.dst_stride + lowx;
		const int32_t *srow = sm + r * PITCH;
		if (l0) orow[lane] = srow[lane];
		if (l1) orow[lane + 32] = srow[lane + 32];
		if (h0) orow[hoff] = srow[TW / 2 + lane];
		if (h1) orow[hoff + 32] = srow[TW / 2 + lane + 32];
	}
}

template<bool REV, int TH>
__global__ void __launch_bounds__(NCOL) dwt_inv_kernel(const DwtPlane *__restrict__ planes, const uint32_t *__restrict__ cta_plane) {
	using G = Geo<REV, TH>;
	constexpr int H = G::H, TW = G::TW, NR = G::NR, HALF = G::HALF, NU = G::NU, PITCH = NCOL + 1;
	__shared__ int32_t sm[NR * PITCH];

	uint32_t cta = blockIdx.x;
	const DwtPlane P = find_plane(planes, cta_plane, cta);
	const int X0 = (cta % P.tiles_x) * TW, Y0 = (cta / P.tiles_x) * TH;
	const int rw = P.rw, rh = P.rh;
	const int t = threadIdx.x;
	const int hpx = 1 - (int) P.cas_x, hpy = 1 - (int) P.cas_y;

	{ // interleave the four sub-bands into the tile (dwt.cpp:1219-1316 for 9/7; implicit in the 5/3 sweeps):
	  // thread = column; LL comes from the previous level's output, the rest from the coefficient plane
		const int gx = reflect(X0 - H + t, rw);
		const bool hx = (gx & 1) == hpx;
		const int sx = (gx >> 1) + (hx ? (int) P.sw : 0);
		const bool interior = Y0 - H >= 0 && Y0 - H + NR <= rh, simple = rh > 2 * H;
		#pragma unroll 8
		for (int r = 0; r < NR; ++r) {
			const int y = Y0 - H + r;
			const int gy = interior ? y : (simple ? reflect1(min(y, rh + H - 1), rh) : reflect(y, rh));
			const bool hy = (gy & 1) == hpy;
			const int sy = (gy >> 1) + (hy ? (int) P.sh : 0);
			const int32_t *p = (!hx && !hy) ? P.src + (size_t) sy * P.src_stride : P.band + (size_t) sy * P.band_stride;
			sm[r * PITCH + t] = p[sx];
		}
	}
	__syncthreads();
	// horizontal first (dwt.cpp:775-803, 1586-1650): work item = half a row, NR rows -> 2*NR items.
	// The two items of a row sit in adjacent lanes; each rewrites only its own valid span, which
	// overlaps the halo the other one reads, hence the warp barrier between the reads and the writes.
	for (int base = 0; base < 2 * NR; base += NCOL) {
		const int item = base + t;
		const bool active = item < 2 * NR;
		const int row = active ? item >> 1 : 0, c0 = (item & 1) * HALF;
		int32_t u[NU];
		#pragma unroll
		for (int i = 0; i < NU; ++i) u[i] = sm[row * PITCH + c0 + i];
		lift_line<false, REV, NU>(u, rw, hpx, (int) P.cas_x);
		__syncwarp();
		if (active) {
			#pragma unroll
			for (int i = H; i < H + HALF; ++i) sm[row * PITCH + c0 + i] = u[i];
		}
	}
	__syncthreads();
	{ // vertical (dwt.cpp:822-853, 1653-1733): thread = column, only the TW valid columns
		int32_t v[NR];
		#pragma unroll
		for (int r = 0; r < NR; ++r) v[r] = sm[r * PITCH + t];
		lift_line<false, REV, NR>(v, rh, hpy, (int) P.cas_y);
		const int gx = X0 + t - H;
		if (t >= H && t < H + TW && gx < rw) {
			int32_t *col = P.dst + gx;
			#pragma unroll
			for (int r = 0; r < TH; ++r) {
				const int gy = Y0 + r;
				if (gy < rh) col[(size_t) gy * P.dst_stride] = v[r + H];
			}
		}
	}
}

template<bool REV>
static void launch_fwd_t(const DwtPlane *p, const uint32_t *m, uint32_t n, int th, cudaStream_t s) {
	if (th == 64) dwt_fwd_kernel<REV, 64><<<n, NCOL, 0, s>>>(p, m);
	else if (th == 32) dwt_fwd_kernel<REV, 32><<<n, NCOL, 0, s>>>(p, m);
	else dwt_fwd_kernel<REV, 16><<<n, NCOL, 0, s>>>(p, m);
}
template<bool REV>
static void launch_inv_t(const DwtPlane *p, const uint32_t *m, uint32_t n, int th, cudaStream_t s) {
	if (th == 64) dwt_inv_kernel<REV, 64><<<n, NCOL, 0, s>>>(p, m);
	else if (th == 32) dwt_inv_kernel<REV, 32><<<n, NCOL, 0, s>>>(p, m);
	else dwt_inv_kernel<REV, 16><<<n, NCOL, 0, s>>>(p, m);
}

// tile_rows > 0: shared-memory kernels above (first generation), 64 / 32 / 16 valid rows per CTA
// tile_rows < 0: streaming kernels of dwt_stream.cuh, -tile_rows rows per work item (one warp each); `total_ctas` counts work items
// Which prefetch queue a kernel uses (dwt_stream.cuh).  Measured on B200 (tools/dwt_bench.py, profiles/README.md): the register
// queue two trips deep wins for all four kernels (configs[1] 9/7: 0.117 / 0.093 ms forward / inverse against 0.127 / 0.103 with the
// shared-memory ring; configs[2] 5/3: 0.42 / 0.40 against 0.46 / 0.56), so it is the default; GB200_DWT_RING=1 selects the ring.  `unroll` (1, 2, 4) selects the depth: ring 4 / 8 / 12 trips, registers 2 / 4 / 6.
static bool use_ring(int reversible) {
	const char *e = getenv("GB200_DWT_RING");
	(void) reversible;
	return e && *e ? atoi(e) != 0 : false;
}

// the prefetch rings want the shared-memory end of the L1 / shared split
template<typename K>
static void prefer_shared(K kernel) { cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); }

// calls f(kernel) with the instantiation for (forward, reversible, ring, unroll)
template<bool FWD, bool REV, int G, bool RING, typename F>
static void with_kernel(F &&f) {
	if (FWD) f(dwt_fwd_stream_kernel<REV, G, RING>);
	else f(dwt_inv_stream_kernel<REV, G, RING>);
}
template<bool FWD, bool REV, typename F>
static void select_kernel(bool ring, int unroll, F &&f) {
	if (ring) {
		if (unroll == 1) with_kernel<FWD, REV, 4, true>(f);
		else if (unroll == 4) with_kernel<FWD, REV, 12, true>(f);
		else with_kernel<FWD, REV, 8, true>(f);
	} else {
		if (unroll == 1) with_kernel<FWD, REV, 2, false>(f);
		else if (unroll == 4) with_kernel<FWD, REV, 6, false>(f);
		else with_kernel<FWD, REV, 4, false>(f);
	}
}
template<typename F>
static void select_kernel(bool fwd, bool rev, bool ring, int unroll, F &&f) {
	if (fwd) { if (rev) select_kernel<true, true>(ring, unroll, f); else select_kernel<true, false>(ring, unroll, f); }
	else { if (rev) select_kernel<false, true>(ring, unroll, f); else select_kernel<false, false>(ring, unroll, f); }
}

static void stream_kernels_setup() {
	static bool done = false; // per process; the attribute is per function
	if (done) return;
	done = true;
	for (int fwd = 0; fwd < 2; ++fwd)
		for (int rev = 0; rev < 2; ++rev)
			for (int unroll : {1, 2, 4}) select_kernel(fwd != 0, rev != 0, true, unroll, [](auto k) { prefer_shared(k); });
}

// launch with programmatic stream serialisation (see dws_launch_dependents / dws_grid_wait in dwt_stream.cuh)
static void launch_stream(bool fwd, bool rev, const DwtPlane *p, const uint32_t *m, uint32_t n, int rows, int unroll, int hl, cudaStream_t s) {
	stream_kernels_setup();
	select_kernel(fwd, rev, use_ring(rev), unroll, [&](auto kernel) {
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((n + DWS_WARPS - 1) / DWS_WARPS); cfg.blockDim = dim3(DWS_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = s;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		cudaLaunchKernelEx(&cfg, kernel, p, m, n, rows, hl);
	});
}

void dwt_stream_shape(int halo_lanes, uint32_t *tw) { *tw = dws_tw(halo_lanes); }

// resident warps (= work items in flight) per SM of the streaming kernel a launch with these parameters uses
int dwt_stream_warps_per_sm(int reversible, int forward, int unroll) {
	stream_kernels_setup();
	int warps = 4 * DWS_WARPS;
	select_kernel(forward != 0, reversible != 0, use_ring(reversible), unroll, [&](auto kernel) {
		int ctas = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, DWS_WARPS * 32, 0) == cudaSuccess && ctas >= 1) warps = ctas * DWS_WARPS;
	});
	return warps;
}

void launch_dwt_fwd(const DwtPlane *planes_dev, const uint32_t *cta_plane_dev, uint32_t total_ctas, int reversible, int tile_rows,
		int unroll, int halo_lanes, cudaStream_t s) {
	if (!total_ctas) return;
	if (tile_rows < 0) {
		launch_stream(true, reversible != 0, planes_dev, cta_plane_dev, total_ctas, -tile_rows, unroll, halo_lanes, s);
	} else if (reversible) launch_fwd_t<true>(planes_dev, cta_plane_dev, total_ctas, tile_rows, s);
	else launch_fwd_t<false>(planes_dev, cta_plane_dev, total_ctas, tile_rows, s);
}

void launch_dwt_inv(const DwtPlane *planes_dev, const uint32_t *cta_plane_dev, uint32_t total_ctas, int reversible, int tile_rows,
		int unroll, int halo_lanes, cudaStream_t s) {
	if (!total_ctas) return;
	if (tile_rows < 0) {
		launch_stream(false, reversible != 0, planes_dev, cta_plane_dev, total_ctas, -tile_rows, unroll, halo_lanes, s);
	} else if (reversible) launch_inv_t<true>(planes_dev, cta_plane_dev, total_ctas, tile_rows, s);
	else launch_inv_t<false>(planes_dev, cta_plane_dev, total_ctas, tile_rows, s);
}

} // namespace gb
