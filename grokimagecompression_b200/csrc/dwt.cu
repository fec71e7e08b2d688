// K2 / K3: forward and inverse wavelet transforms, one decomposition level per launch, every
// lifting step of both directions in one pass over the data.
//
// forward 5/3  WaveletForward.h:40-161 + dwt53.cpp:150-169   (int32, exact)
// forward 9/7  WaveletForward.h:40-161 + dwt97.cpp:90-123    (int32 13-bit fixed point, exact)
// inverse 5/3  dwt.cpp:724-858, 256-363, 661-718             (int32, exact)
// inverse 9/7  dwt.cpp:1544-1738, 1413-1537, constants 172-178 (fp32, multiply-then-add, no FMA)
//
// The kernels are the STREAMING kernels of dwt_stream.cuh (one warp per strip, vertical lifting recurrences in registers,
// horizontal lifting across lanes, no shared memory, no barrier); this file holds their launch code.  (The round-1
// first generation -- a CTA staging a 128-column tile in shared memory -- measured 0.135 / 0.367 ms forward / inverse 9/7
// on configs[1] planes against 0.116 / 0.094 ms and was removed; the numbers are kept in profiles/README.md.)
// Launches are out of place (ping-pong planes): no warp reads what another warp of the launch writes.
//
// Borders use whole-sample symmetric reflection of the source index.  Lifting on a symmetrically
// extended signal keeps it symmetric, so this produces exactly the reference's clamped-neighbour
// results (the (c+c)*x special case of dwt.cpp:1463-1469 rounds like (x+x)*c).  Lines of length 1
// are the exception and are handled explicitly (dwt53.cpp:160, dwt.cpp:344-349, 1482-1490).
#include "common.cuh"
#include "dwt_stream.cuh"

namespace gb {

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

void launch_dwt_fwd(const DwtPlane *planes_dev, const uint32_t *item_plane_dev, uint32_t total_items, int reversible, int rows,
		int unroll, int halo_lanes, cudaStream_t s) {
	if (!total_items) return;
	launch_stream(true, reversible != 0, planes_dev, item_plane_dev, total_items, rows, unroll, halo_lanes, s);
}

void launch_dwt_inv(const DwtPlane *planes_dev, const uint32_t *item_plane_dev, uint32_t total_items, int reversible, int rows,
		int unroll, int halo_lanes, cudaStream_t s) {
	if (!total_items) return;
	launch_stream(false, reversible != 0, planes_dev, item_plane_dev, total_items, rows, unroll, halo_lanes, s);
}

} // namespace gb
