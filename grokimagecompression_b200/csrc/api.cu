// C-ABI layer of libgrok_b200.so (include/grok_b200.h): contexts, plans (geometry + block table +
// device buffers of a tile batch), the encode / decode pipelines and the stage-level entry points.
//
// Geometry mirrors TileComponent::init (TileComponent.cpp:193-489) and the block offsets of
// Tier1::encodeCodeblocks / prepareDecodeCodeblocks (Tier1.cpp:53-86, 117-166).
#include "../../include/grok_b200.h"
#include "common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <cmath>

using namespace gb;

static thread_local std::string g_err = "";

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	g_err = std::string(#call) + ": " + cudaGetErrorString(e_); return GB200_ERR_CUDA; } } while (0)
#define FAIL(code, msg) do { g_err = (msg); return (code); } while (0)

struct gb200_ctx {
	int device;
	cudaStream_t stream;
	uint64_t launches;
};

#ifndef DWT_MIN_ROWS
#define DWT_MIN_ROWS 2   // smallest number of rows per work item the wave model may pick (even)
#endif
static inline uint32_t cdiv2n(uint32_t a, uint32_t n) { return (uint32_t) (((uint64_t) a + ((1ull << n) - 1)) >> n); }
static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

namespace {

struct CompGeom {
	gb200_comp_params p;
	uint32_t w, h;          // size of the stored plane (decoder: the reduced resolution)
	uint32_t stride;
	uint32_t top;           // decomposition level of the plane (0 = full resolution)
	uint64_t plane_off;     // element offset inside the per-component device buffers
	uint32_t levels;        // decompositions present in the plane
};

struct TileGeom {
	uint32_t numcomps, mct, rate_control;
	std::vector<CompGeom> comps;
};

struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
	int alloc(size_t n) {
		if (p) { cudaFree(p); p = nullptr; }
		bytes = n;
		if (!n) return 0;
		return cudaMalloc(&p, n) == cudaSuccess ? 0 : 1;
	}
	void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct LevelLaunch {
	std::vector<DwtPlane> host;
	std::vector<uint32_t> cta_plane; // CTA -> index into host
	DevBuf dev, map;
	uint32_t ctas = 0;
	std::vector<std::pair<uint32_t, uint32_t>> strips; // (strips across, rows incl. the parity offset) of every plane
	int rows = 16;        // rows per work item (one warp each)
	int unroll = 2, halo_lanes = 1;
};

} // namespace

struct gb200_plan {
	gb200_ctx *ctx;
	bool encoder;
	std::vector<TileGeom> tiles;
	uint32_t maxcomps = 0;
	std::vector<gb200_cblk_info> blocks;
	uint64_t pass_slots = 0, samples = 0, data_cap = 0;
	// device planes: one buffer per component index and role (A,B[,C]); tiles back to back
	std::vector<DevBuf> bufA, bufB, bufC;
	std::vector<uint64_t> comp_elems;
	// DWT launches: [level][rev ? 1 : 0]
	std::vector<LevelLaunch> lvl[2];
	uint32_t maxlevels = 0;
	// Tier-1
	std::vector<EncBlock> encblocks;
	std::vector<DecBlock> decblocks;
	DevBuf d_blocks, d_results, d_rates, d_dists, d_scratch, d_data, d_inputs, d_symbols, d_seg_start, d_segs;
	DevBuf d_slopes, d_slope_cache; // allocated by the first gb200_encode_slopes
	bool have_segs = false;
	uint64_t d_data_len = 0;
	uint32_t max_bw = 1, max_bh = 1; // largest code block of the table
	bool uniform = true; // every tile shares mct / qmfbid / shift / range parameters
	bool styles = false; // some component uses code-block style switches
	bool ht = false;     // every component uses the HTJ2K block coder (tcp->isHT is per tile in the reference, T1Factory.cpp:36)
	std::vector<EncResult> h_results;
	// where each tile-component's final decoded plane lives (0 A, 1 B, 2 C)
	std::vector<int> final_role;
	std::vector<DevBuf> stash;
	// narrow-sample boundary: bytes per sample of the host planes (4 = int32, the reference's tile-buffer contract; 1 / 2 =
	// packed image samples, widened / narrowed inside the level-shift + MCT pass)
	uint32_t sample_bytes = 4;
	DevBuf d_total;            // byte count of the compacted code-block data, written by the offsets kernel
	uint64_t *h_total = nullptr; // pinned host copy of it
	cudaEvent_t ev_total = nullptr;
};

extern "C" {

int gb200_abi_version(void) { return GB200_ABI_VERSION; }
const char *gb200_last_error(void) { return g_err.c_str(); }

int gb200_device_count(void) {
	int n = 0;
	return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int gb200_create(int device, gb200_ctx **out) {
	if (!out) FAIL(GB200_ERR_PARAM, "gb200_create: out is NULL");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		FAIL(GB200_ERR_CUDA, std::string("gb200_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
	if (device < 0 || device >= n) FAIL(GB200_ERR_PARAM, "gb200_create: bad device index");
	CK(cudaSetDevice(device));
	gb200_ctx *c = new (std::nothrow) gb200_ctx();
	if (!c) FAIL(GB200_ERR_NOMEM, "out of host memory");
	c->device = device;
	c->launches = 0;
	CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	*out = c;
	return GB200_OK;
}

void gb200_destroy(gb200_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

uint64_t gb200_launch_count(const gb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
void *gb200_stream(const gb200_ctx *ctx) { return ctx ? (void*) ctx->stream : nullptr; }
int gb200_sync(gb200_ctx *ctx) {
	if (!ctx) FAIL(GB200_ERR_PARAM, "ctx is NULL");
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	return GB200_OK;
}

} // extern "C"

// ---- geometry ------------------------------------------------------------------------------------

namespace {

struct BlockGeom {
	uint32_t resno, orient, precno, cblkno, x0, y0, x1, y1, off_x, off_y, band_index;
};

// code blocks of one tile-component in (resno, band, precinct, block) order
static void enumerate_blocks(const gb200_comp_params &p, uint32_t numres_limit, std::vector<BlockGeom> &out) {
	const uint32_t numres = p.numres;
	for (uint32_t resno = 0; resno < numres_limit; ++resno) {
		const uint32_t lvl = numres - 1 - resno;
		const uint32_t rx0 = cdiv2n(p.x0, lvl), ry0 = cdiv2n(p.y0, lvl), rx1 = cdiv2n(p.x1, lvl), ry1 = cdiv2n(p.y1, lvl);
		const uint32_t pdx = p.prcw_expn[resno], pdy = p.prch_expn[resno];
		const uint32_t px0 = (rx0 >> pdx) << pdx, py0 = (ry0 >> pdy) << pdy;
		const uint32_t px1 = cdiv2n(rx1, pdx) << pdx, py1 = cdiv2n(ry1, pdy) << pdy;
		const uint32_t pw = rx0 == rx1 ? 0 : (px1 - px0) >> pdx, ph = ry0 == ry1 ? 0 : (py1 - py0) >> pdy;
		uint32_t gx0, gy0, gwe, ghe, nbands;
		if (resno == 0) { gx0 = px0; gy0 = py0; gwe = pdx; ghe = pdy; nbands = 1; }
		else { gx0 = cdiv2n(px0, 1); gy0 = cdiv2n(py0, 1); gwe = pdx - 1; ghe = pdy - 1; nbands = 3; }
		const uint32_t cwe = std::min(p.cblkw_expn, gwe), che = std::min(p.cblkh_expn, ghe);
		const uint32_t lw = resno ? cdiv2n(p.x1, lvl + 1) - cdiv2n(p.x0, lvl + 1) : 0;
		const uint32_t lh = resno ? cdiv2n(p.y1, lvl + 1) - cdiv2n(p.y0, lvl + 1) : 0;
		for (uint32_t b = 0; b < nbands; ++b) {
			const uint32_t orient = resno == 0 ? 0 : b + 1;
			uint32_t bx0, by0, bx1, by1;
			if (resno == 0) { bx0 = rx0; by0 = ry0; bx1 = rx1; by1 = ry1; }
			else {
				const uint64_t xo = (uint64_t) (orient & 1) << lvl, yo = (uint64_t) (orient >> 1) << lvl;
				const uint64_t rnd = (1ull << (lvl + 1)) - 1;
				bx0 = (uint32_t) (((uint64_t) p.x0 - xo + rnd) >> (lvl + 1));
				by0 = (uint32_t) (((uint64_t) p.y0 - yo + rnd) >> (lvl + 1));
				bx1 = (uint32_t) (((uint64_t) p.x1 - xo + rnd) >> (lvl + 1));
				by1 = (uint32_t) (((uint64_t) p.y1 - yo + rnd) >> (lvl + 1));
			}
			for (uint32_t prc = 0; prc < pw * ph; ++prc) {
				const uint32_t cx0 = gx0 + (prc % pw) * (1u << gwe), cy0 = gy0 + (prc / pw) * (1u << ghe);
				const uint32_t qx0 = std::max(cx0, bx0), qy0 = std::max(cy0, by0);
				const uint32_t qx1 = std::min(cx0 + (1u << gwe), bx1), qy1 = std::min(cy0 + (1u << ghe), by1);
				// an empty band still yields (zero-area) code blocks when its edge is not block aligned
				// (TileComponent.cpp:384-404); they must stay in the table to keep the host's block order
				if (qx1 < qx0 || qy1 < qy0) continue;
				const uint32_t kx0 = (qx0 >> cwe) << cwe, ky0 = (qy0 >> che) << che;
				const uint32_t kx1 = cdiv2n(qx1, cwe) << cwe, ky1 = cdiv2n(qy1, che) << che;
				const uint32_t cw = (kx1 - kx0) >> cwe, ch = (ky1 - ky0) >> che;
				for (uint32_t k = 0; k < cw * ch; ++k) {
					const uint32_t ax = kx0 + (k % cw) * (1u << cwe), ay = ky0 + (k / cw) * (1u << che);
					BlockGeom g;
					g.resno = resno; g.orient = orient; g.precno = prc; g.cblkno = k;
					g.x0 = std::max(ax, qx0); g.y0 = std::max(ay, qy0);
					g.x1 = std::min(ax + (1u << cwe), qx1); g.y1 = std::min(ay + (1u << che), qy1);
					g.off_x = g.x0 - bx0 + ((orient & 1) ? lw : 0);
					g.off_y = g.y0 - by0 + ((orient & 2) ? lh : 0);
					g.band_index = resno == 0 ? 0 : 3 * resno - 2 + b;
					out.push_back(g);
				}
			}
		}
	}
}

static inline int launch_check(gb200_ctx *ctx, int n) {
	ctx->launches += (uint64_t) n;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_err = std::string("kernel launch: ") + cudaGetErrorString(e); return GB200_ERR_CUDA; }
	return GB200_OK;
}

static int32_t *plane_ptr(gb200_plan *pl, int role, uint32_t compno, uint64_t off) {
	DevBuf &b = role == 0 ? pl->bufA[compno] : role == 1 ? pl->bufB[compno] : pl->bufC[compno];
	return reinterpret_cast<int32_t*>(b.p) + off;
}

} // namespace

extern "C" {

void gb200_plan_destroy(gb200_plan *pl) {
	if (!pl) return;
	cudaSetDevice(pl->ctx->device);
	for (auto &b : pl->bufA) b.release();
	for (auto &b : pl->bufB) b.release();
	for (auto &b : pl->bufC) b.release();
	for (auto &b : pl->stash) b.release();
	for (int r = 0; r < 2; ++r) for (auto &l : pl->lvl[r]) { l.dev.release(); l.map.release(); }
	pl->d_blocks.release(); pl->d_results.release(); pl->d_rates.release(); pl->d_dists.release();
	pl->d_slopes.release(); pl->d_slope_cache.release();
	pl->d_scratch.release(); pl->d_data.release(); pl->d_inputs.release(); pl->d_symbols.release(); pl->d_seg_start.release(); pl->d_segs.release();
	pl->d_total.release();
	if (pl->h_total) cudaFreeHost(pl->h_total);
	if (pl->ev_total) cudaEventDestroy(pl->ev_total);
	delete pl;
}

int gb200_plan_create(gb200_ctx *ctx, uint32_t ntiles, const gb200_tile_params *tiles, int is_encoder, gb200_plan **out) {
	if (!ctx || !tiles || !out || !ntiles) FAIL(GB200_ERR_PARAM, "gb200_plan_create: bad arguments");
	CK(cudaSetDevice(ctx->device));
	gb200_plan *pl = new (std::nothrow) gb200_plan();
	if (!pl) FAIL(GB200_ERR_NOMEM, "out of host memory");
	pl->ctx = ctx;
	pl->encoder = is_encoder != 0;
	auto bail = [&](int code, const std::string &msg) { g_err = msg; gb200_plan_destroy(pl); return code; };

	// ---- geometry -------------------------------------------------------------------------------
	for (uint32_t t = 0; t < ntiles; ++t) {
		const gb200_tile_params &tp = tiles[t];
		if (!tp.numcomps || !tp.comps) return bail(GB200_ERR_PARAM, "tile without components");
		if (tp.mct > 1) return bail(GB200_ERR_UNSUPPORTED, "custom (array) MCT is outside the hot path (mct.cpp:429-511)");
		if (tp.mct == 1 && tp.numcomps < 3) return bail(GB200_ERR_PARAM, "MCT needs three components");
		TileGeom tg;
		tg.numcomps = tp.numcomps; tg.mct = tp.mct; tg.rate_control = tp.rate_control;
		pl->maxcomps = std::max(pl->maxcomps, tp.numcomps);
		for (uint32_t c = 0; c < tp.numcomps; ++c) {
			CompGeom cg;
			cg.p = tp.comps[c];
			const gb200_comp_params &p = cg.p;
			if (p.numres < 1 || p.numres > GB200_MAX_RES || p.x1 < p.x0 || p.y1 < p.y0) return bail(GB200_ERR_PARAM, "bad component rectangle / numres");
			if (p.cblk_sty & ~(uint32_t) (STY_ALL | STY_HT)) return bail(GB200_ERR_PARAM, "unknown code-block style bits");
			if (p.cblk_sty & STY_HT) {
				// grk_compress.cpp:1131-1141: HT cannot be combined with another mode switch; it holds for the whole tile
				if (p.cblk_sty != STY_HT) return bail(GB200_ERR_PARAM, "the HT block coder cannot be combined with other code-block styles");
				if (p.roishift) return bail(GB200_ERR_UNSUPPORTED, "ROI up-shift with the HT block coder");
				if ((t || c) && !pl->ht) return bail(GB200_ERR_UNSUPPORTED, "HT and Part-1 block coders mixed in one plan");
				pl->ht = true;
			} else {
				if (pl->ht) return bail(GB200_ERR_UNSUPPORTED, "HT and Part-1 block coders mixed in one plan");
				if (p.cblk_sty) pl->styles = true;
			}
			if (p.roishift > 30 - 1) return bail(GB200_ERR_UNSUPPORTED, "ROI shift of 30 or more bit planes (t1.cpp:1056)");
			if (p.cblkw_expn > 6 || p.cblkh_expn > 6 || p.cblkw_expn < 2 || p.cblkh_expn < 2)
				return bail(GB200_ERR_UNSUPPORTED, "code blocks larger than 64x64 are not implemented");
			uint32_t nd = pl->encoder ? p.numres : (tp.numres_decode ? std::min(tp.numres_decode, p.numres) : p.numres);
			cg.top = p.numres - nd;
			cg.w = cdiv2n(p.x1, cg.top) - cdiv2n(p.x0, cg.top);
			cg.h = cdiv2n(p.y1, cg.top) - cdiv2n(p.y0, cg.top);
			cg.stride = cg.w;
			cg.levels = nd - 1;
			pl->maxlevels = std::max(pl->maxlevels, cg.levels);
			cg.plane_off = 0;
			tg.comps.push_back(cg);
			pl->samples += (uint64_t) cg.w * cg.h;
		}
		if (tp.mct == 1) {
			for (int c = 1; c < 3; ++c)
				if (tg.comps[c].w != tg.comps[0].w || tg.comps[c].h != tg.comps[0].h || tg.comps[c].p.qmfbid != tg.comps[0].p.qmfbid)
					return bail(GB200_ERR_PARAM, "MCT components must share size and wavelet");
		}
		pl->tiles.push_back(std::move(tg));
	}
	// uniform batch: one element-wise launch covers every tile
	{
		const TileGeom &t0 = pl->tiles[0];
		for (auto &tg : pl->tiles) {
			if (tg.numcomps != t0.numcomps || tg.mct != t0.mct) { pl->uniform = false; break; }
			for (uint32_t c = 0; c < tg.numcomps; ++c) {
				const auto &a = tg.comps[c].p, &b = t0.comps[c].p;
				if (a.qmfbid != b.qmfbid || a.dc_shift != b.dc_shift || a.prec != b.prec || a.sgnd != b.sgnd) pl->uniform = false;
			}
		}
	}
	// ---- device planes --------------------------------------------------------------------------
	pl->comp_elems.assign(pl->maxcomps, 0);
	for (auto &tg : pl->tiles)
		for (uint32_t c = 0; c < tg.numcomps; ++c) {
			tg.comps[c].plane_off = pl->comp_elems[c];
			pl->comp_elems[c] += align_up((uint64_t) tg.comps[c].w * tg.comps[c].h, 64);
		}
	if (pl->uniform && pl->tiles[0].mct) // the batched MCT walks the three buffers with one index
		if (pl->comp_elems[0] != pl->comp_elems[1] || pl->comp_elems[0] != pl->comp_elems[2]) pl->uniform = false;
	pl->bufA.resize(pl->maxcomps); pl->bufB.resize(pl->maxcomps); pl->bufC.resize(pl->maxcomps);
	for (uint32_t c = 0; c < pl->maxcomps; ++c) {
		size_t bytes = std::max<uint64_t>(pl->comp_elems[c], 64) * sizeof(int32_t);
		if (pl->bufA[c].alloc(bytes) || pl->bufB[c].alloc(bytes) || (!pl->encoder && pl->bufC[c].alloc(bytes)))
			return bail(GB200_ERR_NOMEM, "cudaMalloc failed for the tile planes");
		cudaMemsetAsync(pl->bufA[c].p, 0, bytes, ctx->stream);
		cudaMemsetAsync(pl->bufB[c].p, 0, bytes, ctx->stream);
		if (!pl->encoder) cudaMemsetAsync(pl->bufC[c].p, 0, bytes, ctx->stream);
	}
	// ---- DWT launch tables ------------------------------------------------------------------------
	for (int r = 0; r < 2; ++r) pl->lvl[r].resize(pl->maxlevels);
	auto level_of = [&](const CompGeom &cg, uint32_t i) { return pl->encoder ? cg.top + i : cg.p.numres - 2 - i; };
	// GB200_DWT_ROWS / GB200_DWT_UNROLL / GB200_DWT_FILL override the tuning below (measurement knobs).
	auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; };
	const int force_rows = env_int("GB200_DWT_ROWS", 0), unroll = env_int("GB200_DWT_UNROLL", 1), fill = env_int("GB200_DWT_FILL", 4);
	const int halo_lanes = env_int("GB200_DWT_HL", 1) == 2 ? 2 : 1;
	uint32_t stw;
	dwt_stream_shape(halo_lanes, &stw);
	for (auto &tg : pl->tiles)
		for (uint32_t c = 0; c < tg.numcomps; ++c) {
			const CompGeom &cg = tg.comps[c];
			for (uint32_t i = 0; i < cg.levels; ++i) {
				const uint32_t lvl = level_of(cg, i);
				const uint32_t rw = cdiv2n(cg.p.x1, lvl) - cdiv2n(cg.p.x0, lvl), rh = cdiv2n(cg.p.y1, lvl) - cdiv2n(cg.p.y0, lvl);
				if (!rw || !rh) continue;
				LevelLaunch &L = pl->lvl[cg.p.qmfbid == 1][i];
				const uint32_t cx = cdiv2n(cg.p.x0, lvl) & 1, cy = cdiv2n(cg.p.y0, lvl) & 1;
				L.strips.emplace_back((rw + cx + stw - 1) / stw, rh + cy);
			}
		}
	{
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
		for (int r = 0; r < 2; ++r)
			for (auto &L : pl->lvl[r]) {
				L.unroll = unroll;
				L.halo_lanes = halo_lanes;
				{
					// Rows per work item.  A launch runs in waves of `slots` resident warps, and a partly filled last wave
					// costs as much as a full one, so the row count is chosen by a small model: waves x trips per item
					// (rows / 2 + the warm-up trips every item spends before its first result + a fixed start-up term).
					// Few items: the shortest strips win (latency of one warp); one wave or a few: the count that just
					// fits; many waves: long strips (least warm-up work).
					const int slots = sms * dwt_stream_warps_per_sm(r, pl->encoder ? 1 : 0, unroll);
					const int warm = (r ? 2 : 4) + fill; // trips: 2 * LAG + start-up (tables, first rows), `fill` is the knob
					int rows = 16;
					double best = 1e300;
					for (int cand = env_int("GB200_DWT_MINROWS", DWT_MIN_ROWS); cand <= 128; cand += cand < 32 ? 2 : 4) {
						uint64_t items = 0;
						for (auto &sp : L.strips) items += (uint64_t) sp.first * ((sp.second + cand - 1) / cand);
						const double w = (double) items / slots;
						const double waves = w <= 4.0 ? std::ceil(w) : w + 0.5;
						const double cost = waves * (cand / 2 + warm);
						if (cost < best * 0.999) { best = cost; rows = cand; }
					}
					if (force_rows >= 2) rows = force_rows & ~1;
					L.rows = rows;
					if (env_int("GB200_DWT_VERBOSE", 0) && !L.strips.empty()) {
						uint64_t items = 0;
						for (auto &sp : L.strips) items += (uint64_t) sp.first * ((sp.second + rows - 1) / rows);
						fprintf(stderr, "[gb200] dwt %s %s level launch %d: %zu planes, %d rows per item, %llu items, %d slots\n", pl->encoder ? "fwd" : "inv",
								r ? "5/3" : "9/7", (int) (&L - &pl->lvl[r][0]), L.strips.size(), rows, (unsigned long long) items, slots);
					}
				}
			}
	}
	pl->final_role.clear();
	for (auto &tg : pl->tiles)
		for (uint32_t c = 0; c < tg.numcomps; ++c) {
			CompGeom &cg = tg.comps[c];
			const gb200_comp_params &p = cg.p;
			const int rev = p.qmfbid == 1;
			int role_final = 0;
			for (uint32_t i = 0; i < cg.levels; ++i) {
				// encoder: i-th launch transforms decomposition level cg.top + i (finest first)
				// decoder: i-th launch reconstructs level (numres-2-i) (coarsest first)
				const uint32_t lvl = level_of(cg, i);
				const uint32_t trows = (uint32_t) pl->lvl[rev][i].rows;
				DwtPlane d;
				memset(&d, 0, sizeof(d));
				d.rw = cdiv2n(p.x1, lvl) - cdiv2n(p.x0, lvl);
				d.rh = cdiv2n(p.y1, lvl) - cdiv2n(p.y0, lvl);
				d.sw = cdiv2n(p.x1, lvl + 1) - cdiv2n(p.x0, lvl + 1);
				d.sh = cdiv2n(p.y1, lvl + 1) - cdiv2n(p.y0, lvl + 1);
				d.cas_x = cdiv2n(p.x0, lvl) & 1;
				d.cas_y = cdiv2n(p.y0, lvl) & 1;
				d.src_stride = d.band_stride = d.dst_stride = cg.stride;
				if (pl->encoder) {
					d.src = plane_ptr(pl, i & 1, c, cg.plane_off);
					d.dst = plane_ptr(pl, (i & 1) ^ 1, c, cg.plane_off);
					role_final = (i & 1) ^ 1;
				} else {
					// LL ping-pongs between B and C; the detail bands always come from A
					int src_role = i == 0 ? 0 : (i & 1 ? 1 : 2);
					int dst_role = i & 1 ? 2 : 1;
					d.src = plane_ptr(pl, src_role, c, cg.plane_off);
					d.band = plane_ptr(pl, 0, c, cg.plane_off);
					d.dst = plane_ptr(pl, dst_role, c, cg.plane_off);
					role_final = dst_role;
				}
				// a strip starts cas columns / rows before the region (a low-pass line comes first)
				d.tiles_x = (d.rw + d.cas_x + stw - 1) / stw;
				d.tiles_y = (d.rh + d.cas_y + trows - 1) / trows;
				if (d.rw == 0 || d.rh == 0) { d.tiles_x = d.tiles_y = 0; }
				LevelLaunch &L = pl->lvl[rev][i];
				d.first_cta = L.ctas;
				L.ctas += d.tiles_x * d.tiles_y;
				if (d.tiles_x * d.tiles_y) {
					L.cta_plane.insert(L.cta_plane.end(), (size_t) d.tiles_x * d.tiles_y, (uint32_t) L.host.size());
					L.host.push_back(d);
				}
			}
			pl->final_role.push_back(role_final);
		}
	for (int r = 0; r < 2; ++r)
		for (auto &L : pl->lvl[r]) {
			if (L.host.empty()) continue;
			if (L.dev.alloc(L.host.size() * sizeof(DwtPlane)) || L.map.alloc(L.cta_plane.size() * sizeof(uint32_t))) return bail(GB200_ERR_NOMEM, "cudaMalloc failed");
			if (cudaMemcpyAsync(L.map.p, L.cta_plane.data(), L.cta_plane.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
				return bail(GB200_ERR_CUDA, "upload of DWT tables failed");
			if (cudaMemcpyAsync(L.dev.p, L.host.data(), L.host.size() * sizeof(DwtPlane), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
				return bail(GB200_ERR_CUDA, "upload of DWT tables failed");
		}
	// ---- block table ------------------------------------------------------------------------------
	std::vector<BlockGeom> geo;
	uint64_t scratch_off = 0, sym_off = 0;
	size_t tc = 0;
	for (uint32_t t = 0; t < pl->tiles.size(); ++t) {
		TileGeom &tg = pl->tiles[t];
		for (uint32_t c = 0; c < tg.numcomps; ++c, ++tc) {
			CompGeom &cg = tg.comps[c];
			const gb200_comp_params &p = cg.p;
			geo.clear();
			enumerate_blocks(p, p.numres - cg.top, geo);
			for (const BlockGeom &g : geo) {
				gb200_cblk_info bi;
				bi.tileno = t; bi.compno = c; bi.resno = g.resno; bi.bandno = g.orient; bi.precno = g.precno; bi.cblkno = g.cblkno;
				bi.x0 = g.x0; bi.y0 = g.y0; bi.x1 = g.x1; bi.y1 = g.y1;
				bi.band_index = g.band_index;
				bi.max_passes = std::max<uint32_t>(1, 3 * std::max<uint32_t>(p.band_numbps[g.band_index], 1) - 2);
				bi.pass_offset = (uint32_t) pl->pass_slots;
				pl->pass_slots += bi.max_passes;
				pl->blocks.push_back(bi);
				const uint32_t bw = g.x1 - g.x0, bh = g.y1 - g.y0;
				pl->max_bw = std::max(pl->max_bw, bw); pl->max_bh = std::max(pl->max_bh, bh);
				if (pl->encoder) {
					// which ping-pong plane holds this sub-band: the output of the launch that produced it
					int role = 0;
					if (cg.levels) {
						uint32_t i = g.resno == 0 ? cg.levels - 1 : p.numres - 1 - g.resno; // launch index
						role = (i & 1) ^ 1;
					}
					EncBlock eb;
					memset(&eb, 0, sizeof(eb));
					eb.src = plane_ptr(pl, role, c, cg.plane_off) + (size_t) g.off_y * cg.stride + g.off_x;
					eb.stride = cg.stride;
					eb.w = (uint16_t) bw; eb.h = (uint16_t) bh;
					eb.orient = (uint8_t) g.orient;
					eb.reversible = p.qmfbid == 1;
					eb.inv_step = (int32_t) p.inv_step[g.band_index];
					eb.pass_offset = bi.pass_offset;
					eb.max_passes = bi.max_passes;
					// the reference's bound (TileProcessor.cpp:2003-2018); terminated passes add a few flush bytes each
					eb.scratch_cap = (uint32_t) align_up((uint64_t) bw * bh * 4 + 2 + (pl->ht ? t1_ht_scratch_extra() : p.cblk_sty ? 4 * bi.max_passes : 0), 16);
					eb.scratch_off = scratch_off;
					eb.sty = (uint8_t) p.cblk_sty;
					eb.band_numbps = (uint8_t) p.band_numbps[g.band_index];
					eb.stepsize = p.stepsize[g.band_index];
					eb.rd_weight = p.rd_weight[g.band_index];
					eb.sym_off = sym_off;
					eb.sym_cap = pl->ht ? 0 : t1_symbol_capacity(bw, bh, std::max<uint32_t>(p.band_numbps[g.band_index], 1)); // no symbol stream in the HT path
					sym_off += eb.sym_cap;
					scratch_off += eb.scratch_cap;
					pl->encblocks.push_back(eb);
				} else {
					DecBlock db;
					memset(&db, 0, sizeof(db));
					db.dst = plane_ptr(pl, 0, c, cg.plane_off) + (size_t) g.off_y * cg.stride + g.off_x;
					db.stride = cg.stride;
					db.w = (uint16_t) bw; db.h = (uint16_t) bh;
					db.orient = (uint8_t) g.orient;
					db.reversible = p.qmfbid == 1;
					db.sty = (uint8_t) p.cblk_sty;
					db.roishift = (uint8_t) p.roishift;
					db.stepsize = p.stepsize[g.band_index];
					db.band_numbps = p.band_numbps[g.band_index];
					pl->decblocks.push_back(db);
				}
			}
		}
	}
	const size_t nb = pl->blocks.size();
	pl->data_cap = scratch_off;
	if (pl->encoder) {
		if (pl->d_blocks.alloc(std::max<size_t>(nb, 1) * sizeof(EncBlock)) || pl->d_results.alloc(std::max<size_t>(nb, 1) * sizeof(EncResult))
				|| pl->d_rates.alloc(std::max<uint64_t>(pl->pass_slots, 1) * sizeof(uint32_t))
				|| pl->d_dists.alloc(std::max<uint64_t>(pl->pass_slots, 1) * sizeof(double))
				|| pl->d_scratch.alloc(std::max<uint64_t>(scratch_off, 16)) || pl->d_data.alloc(std::max<uint64_t>(scratch_off, 16))
				|| pl->d_symbols.alloc(std::max<uint64_t>(sym_off, 16) + 64))
			return bail(GB200_ERR_NOMEM, "cudaMalloc failed for the Tier-1 buffers");
		if (nb && cudaMemcpyAsync(pl->d_blocks.p, pl->encblocks.data(), nb * sizeof(EncBlock), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
			return bail(GB200_ERR_CUDA, "upload of the block table failed");
		cudaMemsetAsync(pl->d_scratch.p, 0, pl->d_scratch.bytes, ctx->stream);
		pl->h_results.resize(nb);
		if (pl->d_total.alloc(sizeof(uint64_t)) || cudaHostAlloc((void**) &pl->h_total, sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess
				|| cudaEventCreateWithFlags(&pl->ev_total, cudaEventDisableTiming) != cudaSuccess)
			return bail(GB200_ERR_NOMEM, "allocation of the byte counter failed");
		cudaMemsetAsync(pl->d_total.p, 0, sizeof(uint64_t), ctx->stream);
	} else {
		if (pl->d_blocks.alloc(std::max<size_t>(nb, 1) * sizeof(DecBlock)) || pl->d_inputs.alloc(std::max<size_t>(nb, 1) * sizeof(DecInput)))
			return bail(GB200_ERR_NOMEM, "cudaMalloc failed for the Tier-1 buffers");
		if (nb && cudaMemcpyAsync(pl->d_blocks.p, pl->decblocks.data(), nb * sizeof(DecBlock), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
			return bail(GB200_ERR_CUDA, "upload of the block table failed");
	}
	if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return bail(GB200_ERR_CUDA, "plan upload failed");
	*out = pl;
	return GB200_OK;
}

uint64_t gb200_enumerate_blocks(const gb200_comp_params *p, uint32_t numres_limit, gb200_cblk_info *out, uint64_t cap) {
	if (!p || p->numres < 1 || p->numres > GB200_MAX_RES) return 0;
	std::vector<BlockGeom> geo;
	enumerate_blocks(*p, numres_limit ? std::min(numres_limit, p->numres) : p->numres, geo);
	uint64_t slots = 0;
	for (size_t i = 0; i < geo.size(); ++i) {
		const BlockGeom &g = geo[i];
		const uint32_t mp = std::max<uint32_t>(1, 3 * std::max<uint32_t>(p->band_numbps[g.band_index], 1) - 2);
		if (out && i < cap) {
			gb200_cblk_info &bi = out[i];
			bi.tileno = 0; bi.compno = 0; bi.resno = g.resno; bi.bandno = g.orient; bi.precno = g.precno; bi.cblkno = g.cblkno;
			bi.x0 = g.x0; bi.y0 = g.y0; bi.x1 = g.x1; bi.y1 = g.y1;
			bi.band_index = g.band_index; bi.pass_offset = (uint32_t) slots; bi.max_passes = mp;
		}
		slots += mp;
	}
	return geo.size();
}

int gb200_precinct_grid(const gb200_comp_params *p, uint32_t resno, uint32_t *pw, uint32_t *ph) {
	if (!p || !pw || !ph || p->numres < 1 || p->numres > GB200_MAX_RES || resno >= p->numres) FAIL(GB200_ERR_PARAM, "gb200_precinct_grid: bad arguments");
	const uint32_t lvl = p->numres - 1 - resno;
	const uint32_t rx0 = cdiv2n(p->x0, lvl), ry0 = cdiv2n(p->y0, lvl), rx1 = cdiv2n(p->x1, lvl), ry1 = cdiv2n(p->y1, lvl);
	const uint32_t pdx = p->prcw_expn[resno], pdy = p->prch_expn[resno];
	*pw = rx0 == rx1 ? 0 : ((cdiv2n(rx1, pdx) << pdx) - ((rx0 >> pdx) << pdx)) >> pdx;
	*ph = ry0 == ry1 ? 0 : ((cdiv2n(ry1, pdy) << pdy) - ((ry0 >> pdy) << pdy)) >> pdy;
	return GB200_OK;
}

uint64_t gb200_plan_num_blocks(const gb200_plan *pl) { return pl ? pl->blocks.size() : 0; }
uint64_t gb200_plan_num_pass_slots(const gb200_plan *pl) { return pl ? pl->pass_slots : 0; }
uint64_t gb200_plan_num_samples(const gb200_plan *pl) { return pl ? pl->samples : 0; }
const gb200_cblk_info *gb200_plan_blocks(const gb200_plan *pl) { return pl ? pl->blocks.data() : nullptr; }
uint64_t gb200_plan_data_capacity(const gb200_plan *pl) { return pl ? pl->data_cap : 0; }

// ---- encode ---------------------------------------------------------------------------------------

int gb200_plan_set_sample_bytes(gb200_plan *pl, uint32_t sample_bytes) {
	if (!pl) FAIL(GB200_ERR_PARAM, "plan is NULL");
	if (sample_bytes != 1 && sample_bytes != 2 && sample_bytes != 4) FAIL(GB200_ERR_PARAM, "sample_bytes must be 1, 2 or 4");
	if (sample_bytes < 4)
		for (auto &tg : pl->tiles) {
			for (auto &cg : tg.comps)
				if (cg.p.prec > 8 * sample_bytes) FAIL(GB200_ERR_PARAM, "component precision does not fit the packed sample type");
			if (tg.mct && (tg.comps[0].p.sgnd != tg.comps[1].p.sgnd || tg.comps[0].p.sgnd != tg.comps[2].p.sgnd))
				FAIL(GB200_ERR_UNSUPPORTED, "packed samples: the three MCT components must share their signedness");
		}
	pl->sample_bytes = sample_bytes;
	return GB200_OK;
}

// packed planes wait in the B buffers (the partner of the first wavelet level's ping-pong, free until then); the widening
// level-shift + MCT pass moves them into the A buffers
static uint8_t *packed_ptr(gb200_plan *pl, int role, uint32_t compno, uint64_t elem_off) {
	DevBuf &b = role == 0 ? pl->bufA[compno] : pl->bufB[compno];
	return reinterpret_cast<uint8_t*>(b.p) + elem_off * pl->sample_bytes;
}

int gb200_encode_upload_packed(gb200_plan *pl, const void *const *planes) {
	if (!pl || !pl->encoder || !planes) FAIL(GB200_ERR_PARAM, "gb200_encode_upload: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	size_t i = 0;
	for (auto &tg : pl->tiles)
		for (uint32_t c = 0; c < tg.numcomps; ++c, ++i) {
			const CompGeom &cg = tg.comps[c];
			if (!planes[i]) FAIL(GB200_ERR_PARAM, "NULL plane pointer");
			size_t bytes = (size_t) cg.w * cg.h * pl->sample_bytes;
			void *dst = pl->sample_bytes == 4 ? (void*) plane_ptr(pl, 0, c, cg.plane_off) : (void*) packed_ptr(pl, 1, c, cg.plane_off);
			if (bytes) CK(cudaMemcpyAsync(dst, planes[i], bytes, cudaMemcpyHostToDevice, pl->ctx->stream));
		}
	return GB200_OK;
}

int gb200_encode_upload(gb200_plan *pl, const int32_t *const *planes) {
	if (pl && pl->sample_bytes != 4) FAIL(GB200_ERR_PARAM, "this plan takes packed samples: use gb200_encode_upload_packed");
	return gb200_encode_upload_packed(pl, reinterpret_cast<const void *const *>(planes));
}

int gb200_encode_stash(gb200_plan *pl) {
	if (!pl || !pl->encoder) FAIL(GB200_ERR_PARAM, "not an encoder plan");
	CK(cudaSetDevice(pl->ctx->device));
	pl->stash.resize(pl->maxcomps);
	for (uint32_t c = 0; c < pl->maxcomps; ++c) { // int32 input lives in A, packed input in the front of B
		const size_t bytes = pl->sample_bytes == 4 ? pl->bufA[c].bytes : pl->bufA[c].bytes / 4 * pl->sample_bytes;
		if (pl->stash[c].bytes != bytes && pl->stash[c].alloc(bytes)) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
		CK(cudaMemcpyAsync(pl->stash[c].p, pl->sample_bytes == 4 ? pl->bufA[c].p : pl->bufB[c].p, bytes, cudaMemcpyDeviceToDevice, pl->ctx->stream));
	}
	return GB200_OK;
}

int gb200_encode_restore(gb200_plan *pl) {
	if (!pl || !pl->encoder || pl->stash.size() != pl->maxcomps) FAIL(GB200_ERR_PARAM, "nothing stashed");
	CK(cudaSetDevice(pl->ctx->device));
	for (uint32_t c = 0; c < pl->maxcomps; ++c)
		CK(cudaMemcpyAsync(pl->sample_bytes == 4 ? pl->bufA[c].p : pl->bufB[c].p, pl->stash[c].p, pl->stash[c].bytes, cudaMemcpyDeviceToDevice, pl->ctx->stream));
	return GB200_OK;
}

// packed input: every component takes the widening pass, also a reversible one without level shift
static int run_dc_mct_fwd_packed(gb200_plan *pl) {
	gb200_ctx *ctx = pl->ctx;
	cudaStream_t s = ctx->stream;
	const uint32_t sb = pl->sample_bytes;
	int n = 0;
	auto one = [&](const TileGeom &tg, bool whole) {
		uint32_t first = 0;
		const auto &p = tg.comps;
		auto off = [&](uint32_t c) { return whole ? (uint64_t) 0 : p[c].plane_off; };
		auto cnt = [&](uint32_t c) { return whole ? pl->comp_elems[c] : (uint64_t) p[c].w * p[c].h; };
		if (tg.mct) {
			launch_mct_fwd_packed(packed_ptr(pl, 1, 0, off(0)), packed_ptr(pl, 1, 1, off(1)), packed_ptr(pl, 1, 2, off(2)),
					plane_ptr(pl, 0, 0, off(0)), plane_ptr(pl, 0, 1, off(1)), plane_ptr(pl, 0, 2, off(2)), cnt(0),
					p[0].p.dc_shift, p[1].p.dc_shift, p[2].p.dc_shift, p[0].p.qmfbid == 1, sb, (int) p[0].p.sgnd, s);
			n++;
			first = 3;
		}
		for (uint32_t c = first; c < tg.numcomps; ++c) {
			launch_dcshift_fwd_packed(packed_ptr(pl, 1, c, off(c)), plane_ptr(pl, 0, c, off(c)), cnt(c), p[c].p.dc_shift, p[c].p.qmfbid == 1, sb,
					(int) p[c].p.sgnd, s);
			n++;
		}
	};
	if (pl->uniform) one(pl->tiles[0], true);
	else for (auto &tg : pl->tiles) one(tg, false);
	return launch_check(ctx, n);
}

static int run_dc_mct_fwd(gb200_plan *pl) {
	if (pl->sample_bytes != 4) return run_dc_mct_fwd_packed(pl);
	gb200_ctx *ctx = pl->ctx;
	cudaStream_t s = ctx->stream;
	int n = 0;
	const TileGeom &t0 = pl->tiles[0];
	if (pl->uniform) {
		uint32_t first = 0;
		if (t0.mct) {
			const auto &p = t0.comps;
			launch_mct_fwd(plane_ptr(pl, 0, 0, 0), plane_ptr(pl, 0, 1, 0), plane_ptr(pl, 0, 2, 0), pl->comp_elems[0],
					p[0].p.dc_shift, p[1].p.dc_shift, p[2].p.dc_shift, p[0].p.qmfbid == 1, 1, s);
			n++;
			first = 3;
		}
		for (uint32_t c = first; c < t0.numcomps; ++c) {
			const auto &p = t0.comps[c].p;
			if (p.qmfbid == 1 && p.dc_shift == 0) continue;
			launch_dcshift_fwd(plane_ptr(pl, 0, c, 0), pl->comp_elems[c], p.dc_shift, p.qmfbid == 1, s);
			n++;
		}
	} else {
		for (auto &tg : pl->tiles) {
			uint32_t first = 0;
			if (tg.mct) {
				const auto &p = tg.comps;
				launch_mct_fwd(plane_ptr(pl, 0, 0, p[0].plane_off), plane_ptr(pl, 0, 1, p[1].plane_off), plane_ptr(pl, 0, 2, p[2].plane_off),
						(uint64_t) p[0].w * p[0].h, p[0].p.dc_shift, p[1].p.dc_shift, p[2].p.dc_shift, p[0].p.qmfbid == 1, 1, s);
				n++;
				first = 3;
			}
			for (uint32_t c = first; c < tg.numcomps; ++c) {
				const auto &cg = tg.comps[c];
				if (cg.p.qmfbid == 1 && cg.p.dc_shift == 0) continue;
				launch_dcshift_fwd(plane_ptr(pl, 0, c, cg.plane_off), (uint64_t) cg.w * cg.h, cg.p.dc_shift, cg.p.qmfbid == 1, s);
				n++;
			}
		}
	}
	return launch_check(ctx, n);
}

static int run_dwt(gb200_plan *pl, bool fwd) {
	gb200_ctx *ctx = pl->ctx;
	int n = 0;
	const char *only_env = getenv("GB200_DWT_ONLY"); // measurement knob (tools/dwt_bench.py): run a single level launch
	const int only = only_env && *only_env ? atoi(only_env) : -1;
	for (uint32_t i = 0; i < pl->maxlevels; ++i)
		for (int r = 0; r < 2; ++r) {
			LevelLaunch &L = pl->lvl[r][i];
			if (!L.ctas || (only >= 0 && (int) i != only)) continue;
			if (fwd) launch_dwt_fwd((const DwtPlane*) L.dev.p, (const uint32_t*) L.map.p, L.ctas, r, L.rows, L.unroll, L.halo_lanes, ctx->stream);
			else launch_dwt_inv((const DwtPlane*) L.dev.p, (const uint32_t*) L.map.p, L.ctas, r, L.rows, L.unroll, L.halo_lanes, ctx->stream);
			n++;
		}
	return launch_check(ctx, n);
}

static int run_t1_enc(gb200_plan *pl) {
	gb200_ctx *ctx = pl->ctx;
	const uint32_t nb = (uint32_t) pl->blocks.size();
	if (!nb) return GB200_OK;
	int rc = 0;
	for (auto &tg : pl->tiles) rc |= (int) tg.rate_control;
	if (pl->ht)
		launch_t1_ht_encode((const EncBlock*) pl->d_blocks.p, nb, (uint8_t*) pl->d_scratch.p, (EncResult*) pl->d_results.p,
				(uint32_t*) pl->d_rates.p, (double*) pl->d_dists.p, ctx->stream);
	else
		launch_t1_encode((const EncBlock*) pl->d_blocks.p, nb, rc, pl->styles ? 1 : 0, (uint8_t*) pl->d_symbols.p, (uint8_t*) pl->d_scratch.p,
				(EncResult*) pl->d_results.p, (uint32_t*) pl->d_rates.p, (double*) pl->d_dists.p, ctx->stream);
	launch_t1_gather((const EncBlock*) pl->d_blocks.p, (EncResult*) pl->d_results.p, nb, (const uint8_t*) pl->d_scratch.p,
			(uint8_t*) pl->d_data.p, (uint64_t*) pl->d_total.p, ctx->stream);
	return launch_check(ctx, pl->ht ? 3 : 4);
}

int gb200_encode_run_stage(gb200_plan *pl, int stage) {
	if (!pl || !pl->encoder) FAIL(GB200_ERR_PARAM, "not an encoder plan");
	CK(cudaSetDevice(pl->ctx->device));
	if (stage == 0) return run_dc_mct_fwd(pl);
	if (stage == 1) return run_dwt(pl, true);
	if (stage == 2) return run_t1_enc(pl);
	FAIL(GB200_ERR_PARAM, "stage must be 0, 1 or 2");
}

int gb200_encode_run(gb200_plan *pl) {
	for (int s = 0; s < 3; ++s) { int rc = gb200_encode_run_stage(pl, s); if (rc) return rc; }
	return GB200_OK;
}

int gb200_encode_download(gb200_plan *pl, gb200_cblk_enc *blocks, uint32_t *rates, double *dists, uint8_t *data,
		uint64_t data_capacity, uint64_t *data_len) {
	if (!pl || !pl->encoder || !blocks || !rates || !dists || !data_len) FAIL(GB200_ERR_PARAM, "gb200_encode_download: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	cudaStream_t s = pl->ctx->stream;
	const size_t nb = pl->blocks.size();
	static_assert(sizeof(gb200_cblk_enc) == sizeof(EncResult), "ABI mismatch");
	// One pass over PCIe: the byte count of the compacted data comes back first (8 bytes behind the kernels), and while
	// the block results and pass tables are still in flight the data copy is queued right behind them.
	uint64_t total = 0;
	if (nb) {
		CK(cudaMemcpyAsync(pl->h_total, pl->d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
		CK(cudaEventRecord(pl->ev_total, s));
		CK(cudaMemcpyAsync(blocks, pl->d_results.p, nb * sizeof(EncResult), cudaMemcpyDeviceToHost, s));
		CK(cudaMemcpyAsync(rates, pl->d_rates.p, pl->pass_slots * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		bool rc = false;
		for (auto &tg : pl->tiles) rc |= tg.rate_control != 0;
		if (rc) CK(cudaMemcpyAsync(dists, pl->d_dists.p, pl->pass_slots * sizeof(double), cudaMemcpyDeviceToHost, s));
		else memset(dists, 0, pl->pass_slots * sizeof(double)); // no distortions are computed without rate control
		CK(cudaEventSynchronize(pl->ev_total));
		total = *pl->h_total;
		if (total <= data_capacity && total && data) CK(cudaMemcpyAsync(data, pl->d_data.p, total, cudaMemcpyDeviceToHost, s));
	}
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	for (size_t i = 0; i < nb; ++i)
		if (blocks[i].numpasses == 0xFFFFFFFFu) FAIL(GB200_ERR_CAPACITY, "a code block overflowed its byte or pass budget");
	*data_len = total;
	if (total > data_capacity || (total && !data)) FAIL(GB200_ERR_CAPACITY, "data buffer too small for the compressed code blocks");
	return GB200_OK;
}

// PCRD preparation: RateControl::convexHull for every block of the last encode run, on the device
int gb200_encode_slopes(gb200_plan *pl, uint16_t *slopes) {
	if (!pl || !pl->encoder || !slopes) FAIL(GB200_ERR_PARAM, "gb200_encode_slopes: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	cudaStream_t s = pl->ctx->stream;
	const uint64_t slots = std::max<uint64_t>(pl->pass_slots, 1);
	if (!pl->d_slopes.p && (pl->d_slopes.alloc(slots * sizeof(uint16_t)) || pl->d_slope_cache.alloc(slots * sizeof(double))))
		FAIL(GB200_ERR_NOMEM, "cudaMalloc failed for the slope tables");
	CK(cudaMemsetAsync(pl->d_slopes.p, 0, slots * sizeof(uint16_t), s));
	launch_rd_slopes((const EncBlock*) pl->d_blocks.p, (const EncResult*) pl->d_results.p, (uint32_t) pl->blocks.size(),
			(const uint32_t*) pl->d_rates.p, (const double*) pl->d_dists.p, (uint16_t*) pl->d_slopes.p, (double*) pl->d_slope_cache.p, s);
	int rc = launch_check(pl->ctx, 1);
	if (rc) return rc;
	if (pl->pass_slots) CK(cudaMemcpyAsync(slopes, pl->d_slopes.p, pl->pass_slots * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	return GB200_OK;
}

int gb200_encode_tiles(gb200_plan *pl, const int32_t *const *planes, gb200_cblk_enc *blocks, uint32_t *rates, double *dists,
		uint8_t *data, uint64_t data_capacity, uint64_t *data_len) {
	int rc = gb200_encode_upload(pl, planes);
	if (rc) return rc;
	rc = gb200_encode_run(pl);
	if (rc) return rc;
	return gb200_encode_download(pl, blocks, rates, dists, data, data_capacity, data_len);
}

int gb200_encode_tiles_packed(gb200_plan *pl, const void *const *planes, gb200_cblk_enc *blocks, uint32_t *rates, double *dists,
		uint8_t *data, uint64_t data_capacity, uint64_t *data_len) {
	int rc = gb200_encode_upload_packed(pl, planes);
	if (rc) return rc;
	rc = gb200_encode_run(pl);
	if (rc) return rc;
	return gb200_encode_download(pl, blocks, rates, dists, data, data_capacity, data_len);
}

int gb200_encode_get_coefficients(gb200_plan *pl, uint32_t tileno, uint32_t compno, int32_t *out) {
	if (!pl || !pl->encoder || !out || tileno >= pl->tiles.size() || compno >= pl->tiles[tileno].numcomps)
		FAIL(GB200_ERR_PARAM, "gb200_encode_get_coefficients: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	cudaStream_t s = pl->ctx->stream;
	const CompGeom &cg = pl->tiles[tileno].comps[compno];
	const gb200_comp_params &p = cg.p;
	if (!cg.w || !cg.h) return GB200_OK;
	// assemble the Mallat layout from the two ping-pong planes: launch i wrote level i to plane (i&1)^1
	if (cg.levels == 0)
		CK(cudaMemcpyAsync(out, plane_ptr(pl, 0, compno, cg.plane_off), (size_t) cg.w * cg.h * 4, cudaMemcpyDeviceToHost, s));
	for (uint32_t i = 0; i < cg.levels; ++i) {
		const uint32_t rw = cdiv2n(p.x1, i) - cdiv2n(p.x0, i), rh = cdiv2n(p.y1, i) - cdiv2n(p.y0, i);
		if (!rw || !rh) continue;
		const int32_t *src = plane_ptr(pl, (i & 1) ^ 1, compno, cg.plane_off);
		CK(cudaMemcpy2DAsync(out, (size_t) cg.stride * 4, src, (size_t) cg.stride * 4, (size_t) rw * 4, rh, cudaMemcpyDeviceToHost, s));
	}
	CK(cudaStreamSynchronize(s));
	return GB200_OK;
}

// ---- decode ---------------------------------------------------------------------------------------

int gb200_decode_upload(gb200_plan *pl, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len) {
	if (!pl || pl->encoder || !blocks) FAIL(GB200_ERR_PARAM, "gb200_decode_upload: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	cudaStream_t s = pl->ctx->stream;
	const size_t nb = pl->blocks.size();
	static_assert(sizeof(gb200_cblk_dec) == sizeof(DecInput), "ABI mismatch");
	for (size_t i = 0; i < nb; ++i) {
		if (blocks[i].data_len && blocks[i].data_offset + blocks[i].data_len > data_len)
			FAIL(GB200_ERR_PARAM, "code block bytes lie outside the data buffer");
		// the reference fails the decode of such a block (t1.cpp:1055-1058); never leave it silently zero
		if (blocks[i].data_len && blocks[i].numbps + pl->decblocks[i].roishift > 30)
			FAIL(GB200_ERR_UNSUPPORTED, "a code block has more than 30 bit planes (t1.cpp:1056)");
		if (pl->ht && blocks[i].data_len) {
			if (blocks[i].numpasses > 1) FAIL(GB200_ERR_UNSUPPORTED, "HT blocks with refinement passes (SigProp / MagRef) are not implemented: cleanup pass only");
			if (blocks[i].numbps > pl->decblocks[i].band_numbps) FAIL(GB200_ERR_PARAM, "HT block with more bit planes than its band");
		}
	}
	if (pl->d_data.bytes < data_len + T1_DEC_DATA_SLACK) {
		if (pl->d_data.alloc(align_up(data_len + T1_DEC_DATA_SLACK, 256))) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed for the compressed data");
	}
	pl->d_data_len = data_len;
	if (nb) CK(cudaMemcpyAsync(pl->d_inputs.p, blocks, nb * sizeof(DecInput), cudaMemcpyHostToDevice, s));
	if (data_len) CK(cudaMemcpyAsync(pl->d_data.p, data, data_len, cudaMemcpyHostToDevice, s));
	return GB200_OK;
}

int gb200_decode_set_segments(gb200_plan *pl, const uint32_t *seg_start, const gb200_cblk_seg *segs) {
	if (!pl || pl->encoder) FAIL(GB200_ERR_PARAM, "not a decoder plan");
	CK(cudaSetDevice(pl->ctx->device));
	static_assert(sizeof(gb200_cblk_seg) == sizeof(DecSeg), "ABI mismatch");
	if (!seg_start || !segs) { pl->have_segs = false; return GB200_OK; }
	const size_t nb = pl->blocks.size();
	for (size_t i = 0; i < nb; ++i)
		if (seg_start[i + 1] < seg_start[i]) FAIL(GB200_ERR_PARAM, "seg_start must be non-decreasing");
	const size_t ns = seg_start[nb];
	if (pl->d_seg_start.bytes < (nb + 1) * 4 && pl->d_seg_start.alloc((nb + 1) * 4)) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
	if (pl->d_segs.bytes < std::max<size_t>(ns, 1) * sizeof(DecSeg) && pl->d_segs.alloc(std::max<size_t>(ns, 1) * sizeof(DecSeg))) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
	CK(cudaMemcpyAsync(pl->d_seg_start.p, seg_start, (nb + 1) * 4, cudaMemcpyHostToDevice, pl->ctx->stream));
	if (ns) CK(cudaMemcpyAsync(pl->d_segs.p, segs, ns * sizeof(DecSeg), cudaMemcpyHostToDevice, pl->ctx->stream));
	CK(cudaStreamSynchronize(pl->ctx->stream)); /* the caller's arrays may go away */
	pl->have_segs = true;
	return GB200_OK;
}

static int run_t1_dec(gb200_plan *pl) {
	const uint32_t nb = (uint32_t) pl->blocks.size();
	if (!nb) return GB200_OK;
	if (pl->ht) {
		// (a segment table, if the caller set one, holds one segment per block here: gb200_decode_upload refuses HT blocks with more passes)
		launch_t1_ht_decode((const DecBlock*) pl->d_blocks.p, (const DecInput*) pl->d_inputs.p, nb, (const uint8_t*) pl->d_data.p, pl->ctx->stream);
		return launch_check(pl->ctx, 1);
	}
	if (launch_t1_decode((const DecBlock*) pl->d_blocks.p, (const DecInput*) pl->d_inputs.p, nb, (const uint8_t*) pl->d_data.p,
			pl->max_bw, pl->max_bh, pl->styles ? 1 : 0, pl->have_segs ? (const uint32_t*) pl->d_seg_start.p : nullptr,
			pl->have_segs ? (const DecSeg*) pl->d_segs.p : nullptr, pl->ctx->stream))
		FAIL(GB200_ERR_UNSUPPORTED, "Tier-1 decode: code block state does not fit in shared memory");
	return launch_check(pl->ctx, T1_DEC_LAUNCHES);
}

static void sample_range(const gb200_comp_params &p, int32_t &lo, int32_t &hi) {
	if (p.sgnd) { lo = -(1 << (p.prec - 1)); hi = (1 << (p.prec - 1)) - 1; }
	else { lo = 0; hi = (int32_t) ((1u << p.prec) - 1); }
}

// packed output: the finished planes are narrowed into the A buffers (the coefficient planes, consumed by then)
static int run_mct_dc_inv_packed(gb200_plan *pl) {
	gb200_ctx *ctx = pl->ctx;
	cudaStream_t s = ctx->stream;
	const uint32_t sb = pl->sample_bytes;
	int n = 0;
	bool same_role = true;
	for (int r : pl->final_role) if (r != pl->final_role[0]) same_role = false;
	size_t tc = 0;
	auto one = [&](const TileGeom &tg, bool whole) {
		uint32_t first = 0;
		const auto &p = tg.comps;
		auto off = [&](uint32_t c) { return whole ? (uint64_t) 0 : p[c].plane_off; };
		auto cnt = [&](uint32_t c) { return whole ? pl->comp_elems[c] : (uint64_t) p[c].w * p[c].h; };
		auto role = [&](uint32_t c) { return pl->final_role[whole ? 0 : tc + c]; };
		if (tg.mct) {
			int32_t sh[3], lo[3], hi[3];
			for (int c = 0; c < 3; ++c) { sh[c] = p[c].p.dc_shift; sample_range(p[c].p, lo[c], hi[c]); }
			launch_mct_inv_packed(plane_ptr(pl, role(0), 0, off(0)), plane_ptr(pl, role(1), 1, off(1)), plane_ptr(pl, role(2), 2, off(2)),
					packed_ptr(pl, 0, 0, off(0)), packed_ptr(pl, 0, 1, off(1)), packed_ptr(pl, 0, 2, off(2)), cnt(0), sh, lo, hi,
					p[0].p.qmfbid == 1, sb, (int) p[0].p.sgnd, s);
			n++;
			first = 3;
		}
		for (uint32_t c = first; c < tg.numcomps; ++c) {
			int32_t lo, hi;
			sample_range(p[c].p, lo, hi);
			launch_dcshift_inv_packed(plane_ptr(pl, role(c), c, off(c)), packed_ptr(pl, 0, c, off(c)), cnt(c), p[c].p.dc_shift, p[c].p.qmfbid == 1,
					lo, hi, sb, (int) p[c].p.sgnd, s);
			n++;
		}
		tc += tg.numcomps;
	};
	if (pl->uniform && same_role) one(pl->tiles[0], true);
	else for (auto &tg : pl->tiles) one(tg, false);
	return launch_check(ctx, n);
}

static int run_mct_dc_inv(gb200_plan *pl) {
	if (pl->sample_bytes != 4) return run_mct_dc_inv_packed(pl);
	gb200_ctx *ctx = pl->ctx;
	cudaStream_t s = ctx->stream;
	int n = 0;
	size_t tc = 0;
	auto range = [](const gb200_comp_params &p, int32_t &lo, int32_t &hi) { sample_range(p, lo, hi); };
	const TileGeom &t0 = pl->tiles[0];
	bool same_role = true;
	for (int r : pl->final_role) if (r != pl->final_role[0]) same_role = false;
	if (pl->uniform && same_role) {
		const int role = pl->final_role[0];
		uint32_t first = 0;
		if (t0.mct) {
			int32_t sh[3], lo[3], hi[3];
			for (int c = 0; c < 3; ++c) { sh[c] = t0.comps[c].p.dc_shift; range(t0.comps[c].p, lo[c], hi[c]); }
			launch_mct_inv(plane_ptr(pl, role, 0, 0), plane_ptr(pl, role, 1, 0), plane_ptr(pl, role, 2, 0), pl->comp_elems[0], sh, lo, hi,
					t0.comps[0].p.qmfbid == 1, 1, s);
			n++;
			first = 3;
		}
		for (uint32_t c = first; c < t0.numcomps; ++c) {
			int32_t lo, hi;
			range(t0.comps[c].p, lo, hi);
			launch_dcshift_inv(plane_ptr(pl, role, c, 0), pl->comp_elems[c], t0.comps[c].p.dc_shift, t0.comps[c].p.qmfbid == 1, lo, hi, s);
			n++;
		}
	} else {
		for (auto &tg : pl->tiles) {
			uint32_t first = 0;
			if (tg.mct) {
				int32_t sh[3], lo[3], hi[3];
				for (int c = 0; c < 3; ++c) { sh[c] = tg.comps[c].p.dc_shift; range(tg.comps[c].p, lo[c], hi[c]); }
				const int role = pl->final_role[tc];
				launch_mct_inv(plane_ptr(pl, role, 0, tg.comps[0].plane_off), plane_ptr(pl, pl->final_role[tc + 1], 1, tg.comps[1].plane_off),
						plane_ptr(pl, pl->final_role[tc + 2], 2, tg.comps[2].plane_off), (uint64_t) tg.comps[0].w * tg.comps[0].h, sh, lo, hi,
						tg.comps[0].p.qmfbid == 1, 1, s);
				n++;
				first = 3;
			}
			for (uint32_t c = first; c < tg.numcomps; ++c) {
				const auto &cg = tg.comps[c];
				int32_t lo, hi;
				range(cg.p, lo, hi);
				launch_dcshift_inv(plane_ptr(pl, pl->final_role[tc + c], c, cg.plane_off), (uint64_t) cg.w * cg.h, cg.p.dc_shift,
						cg.p.qmfbid == 1, lo, hi, s);
				n++;
			}
			tc += tg.numcomps;
		}
	}
	return launch_check(ctx, n);
}

int gb200_decode_run_stage(gb200_plan *pl, int stage) {
	if (!pl || pl->encoder) FAIL(GB200_ERR_PARAM, "not a decoder plan");
	CK(cudaSetDevice(pl->ctx->device));
	if (stage == 0) return run_mct_dc_inv(pl);
	if (stage == 1) return run_dwt(pl, false);
	if (stage == 2) return run_t1_dec(pl);
	FAIL(GB200_ERR_PARAM, "stage must be 0, 1 or 2");
}

int gb200_decode_run(gb200_plan *pl) {
	for (int s = 2; s >= 0; --s) { int rc = gb200_decode_run_stage(pl, s); if (rc) return rc; }
	return GB200_OK;
}

int gb200_decode_download_packed(gb200_plan *pl, void *const *planes_out) {
	if (!pl || pl->encoder || !planes_out) FAIL(GB200_ERR_PARAM, "gb200_decode_download: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	cudaStream_t s = pl->ctx->stream;
	size_t i = 0;
	for (auto &tg : pl->tiles)
		for (uint32_t c = 0; c < tg.numcomps; ++c, ++i) {
			const CompGeom &cg = tg.comps[c];
			size_t bytes = (size_t) cg.w * cg.h * pl->sample_bytes;
			const void *src = pl->sample_bytes == 4 ? (const void*) plane_ptr(pl, pl->final_role[i], c, cg.plane_off) : (const void*) packed_ptr(pl, 0, c, cg.plane_off);
			if (bytes) CK(cudaMemcpyAsync(planes_out[i], src, bytes, cudaMemcpyDeviceToHost, s));
		}
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	return GB200_OK;
}

int gb200_decode_download(gb200_plan *pl, int32_t *const *planes_out) {
	if (pl && pl->sample_bytes != 4) FAIL(GB200_ERR_PARAM, "this plan returns packed samples: use gb200_decode_download_packed");
	return gb200_decode_download_packed(pl, reinterpret_cast<void *const *>(planes_out));
}

int gb200_decode_tiles(gb200_plan *pl, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len, int32_t *const *planes_out) {
	int rc = gb200_decode_upload(pl, blocks, data, data_len);
	if (rc) return rc;
	rc = gb200_decode_run(pl);
	if (rc) return rc;
	return gb200_decode_download(pl, planes_out);
}

int gb200_decode_tiles_packed(gb200_plan *pl, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len, void *const *planes_out) {
	int rc = gb200_decode_upload(pl, blocks, data, data_len);
	if (rc) return rc;
	rc = gb200_decode_run(pl);
	if (rc) return rc;
	return gb200_decode_download_packed(pl, planes_out);
}

int gb200_decode_set_coefficients(gb200_plan *pl, uint32_t tileno, uint32_t compno, const int32_t *in) {
	if (!pl || pl->encoder || !in || tileno >= pl->tiles.size() || compno >= pl->tiles[tileno].numcomps)
		FAIL(GB200_ERR_PARAM, "gb200_decode_set_coefficients: bad arguments");
	CK(cudaSetDevice(pl->ctx->device));
	const CompGeom &cg = pl->tiles[tileno].comps[compno];
	size_t bytes = (size_t) cg.w * cg.h * sizeof(int32_t);
	if (bytes) CK(cudaMemcpyAsync(plane_ptr(pl, 0, compno, cg.plane_off), in, bytes, cudaMemcpyHostToDevice, pl->ctx->stream));
	return GB200_OK;
}

// ---- stage-level entry points ---------------------------------------------------------------------

static int staged3(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n, int which) {
	if (!ctx || !c0 || !c1 || !c2) FAIL(GB200_ERR_PARAM, "bad arguments");
	if (!n) return GB200_OK;
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	DevBuf d[3];
	int32_t *h[3] = {c0, c1, c2};
	for (int i = 0; i < 3; ++i) {
		if (d[i].alloc(n * 4)) { for (auto &b : d) b.release(); FAIL(GB200_ERR_NOMEM, "cudaMalloc failed"); }
		cudaMemcpyAsync(d[i].p, h[i], n * 4, cudaMemcpyHostToDevice, s);
	}
	int32_t *p0 = (int32_t*) d[0].p, *p1 = (int32_t*) d[1].p, *p2 = (int32_t*) d[2].p;
	if (which == 0) launch_mct_fwd(p0, p1, p2, n, 0, 0, 0, 1, 0, s);
	else if (which == 1) launch_mct_inv(p0, p1, p2, n, nullptr, nullptr, nullptr, 1, 0, s);
	else if (which == 2) launch_mct_fwd(p0, p1, p2, n, 0, 0, 0, 0, 0, s);
	else launch_mct_inv(p0, p1, p2, n, nullptr, nullptr, nullptr, 0, 0, s);
	int rc = launch_check(ctx, 1);
	for (int i = 0; i < 3; ++i) cudaMemcpyAsync(h[i], d[i].p, n * 4, cudaMemcpyDeviceToHost, s);
	cudaError_t e = cudaStreamSynchronize(s);
	for (auto &b : d) b.release();
	if (rc) return rc;
	if (e != cudaSuccess) FAIL(GB200_ERR_CUDA, cudaGetErrorString(e));
	return GB200_OK;
}

int gb200_mct_encode_rev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { return staged3(ctx, c0, c1, c2, n, 0); }
int gb200_mct_decode_rev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { return staged3(ctx, c0, c1, c2, n, 1); }
int gb200_mct_encode_irrev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { return staged3(ctx, c0, c1, c2, n, 2); }
int gb200_mct_decode_irrev(gb200_ctx *ctx, float *c0, float *c1, float *c2, uint64_t n) {
	return staged3(ctx, (int32_t*) c0, (int32_t*) c1, (int32_t*) c2, n, 3);
}

static int staged1(gb200_ctx *ctx, int32_t *x, uint64_t n, int32_t shift, int qmfbid, bool fwd, int32_t lo, int32_t hi) {
	if (!ctx || !x) FAIL(GB200_ERR_PARAM, "bad arguments");
	if (!n) return GB200_OK;
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	DevBuf d;
	if (d.alloc(n * 4)) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
	cudaMemcpyAsync(d.p, x, n * 4, cudaMemcpyHostToDevice, s);
	if (fwd) launch_dcshift_fwd((int32_t*) d.p, n, shift, qmfbid == 1, s);
	else launch_dcshift_inv((int32_t*) d.p, n, shift, qmfbid == 1, lo, hi, s);
	int rc = launch_check(ctx, 1);
	cudaMemcpyAsync(x, d.p, n * 4, cudaMemcpyDeviceToHost, s);
	cudaError_t e = cudaStreamSynchronize(s);
	d.release();
	if (rc) return rc;
	if (e != cudaSuccess) FAIL(GB200_ERR_CUDA, cudaGetErrorString(e));
	return GB200_OK;
}

int gb200_dc_shift_encode(gb200_ctx *ctx, int32_t *x, uint64_t n, int32_t shift, int qmfbid) {
	return staged1(ctx, x, n, shift, qmfbid, true, 0, 0);
}
int gb200_dc_shift_decode(gb200_ctx *ctx, int32_t *x, uint64_t n, int32_t shift, int qmfbid, int32_t lo, int32_t hi) {
	return staged1(ctx, x, n, shift, qmfbid, false, lo, hi);
}

static void fill_comp(gb200_comp_params &p, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres, int qmfbid) {
	memset(&p, 0, sizeof(p));
	p.x0 = x0; p.y0 = y0; p.x1 = x1; p.y1 = y1; p.numres = numres;
	p.cblkw_expn = p.cblkh_expn = 6;
	for (int i = 0; i < GB200_MAX_RES; ++i) p.prcw_expn[i] = p.prch_expn[i] = 15;
	p.qmfbid = (uint32_t) qmfbid; p.prec = 8;
	for (int i = 0; i < GB200_MAX_BANDS; ++i) { p.stepsize[i] = 1.0f; p.inv_step[i] = 8192; p.band_numbps[i] = 1; p.rd_weight[i] = 1.0; }
}

int gb200_dwt_encode(gb200_ctx *ctx, int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres, int qmfbid) {
	if (!ctx || !buf) FAIL(GB200_ERR_PARAM, "bad arguments");
	gb200_comp_params cp;
	fill_comp(cp, x0, y0, x1, y1, numres, qmfbid);
	gb200_tile_params tp;
	memset(&tp, 0, sizeof(tp));
	tp.numcomps = 1; tp.comps = &cp;
	gb200_plan *pl = nullptr;
	int rc = gb200_plan_create(ctx, 1, &tp, 1, &pl);
	if (rc) return rc;
	const int32_t *planes[1] = {buf};
	rc = gb200_encode_upload(pl, planes);
	if (!rc) rc = gb200_encode_run_stage(pl, 1);
	if (!rc) rc = gb200_encode_get_coefficients(pl, 0, 0, buf);
	gb200_plan_destroy(pl);
	return rc;
}

int gb200_dwt_decode(gb200_ctx *ctx, int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
		uint32_t numres_decode, int qmfbid) {
	if (!ctx || !buf) FAIL(GB200_ERR_PARAM, "bad arguments");
	gb200_comp_params cp;
	fill_comp(cp, x0, y0, x1, y1, numres, qmfbid);
	gb200_tile_params tp;
	memset(&tp, 0, sizeof(tp));
	tp.numcomps = 1; tp.comps = &cp; tp.numres_decode = numres_decode;
	gb200_plan *pl = nullptr;
	int rc = gb200_plan_create(ctx, 1, &tp, 0, &pl);
	if (rc) return rc;
	rc = gb200_decode_set_coefficients(pl, 0, 0, buf);
	if (!rc) rc = gb200_decode_run_stage(pl, 1);
	if (!rc) {
		const CompGeom &cg = pl->tiles[0].comps[0];
		size_t bytes = (size_t) cg.w * cg.h * 4;
		if (bytes && cudaMemcpyAsync(buf, plane_ptr(pl, pl->final_role[0], 0, cg.plane_off), bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = GB200_ERR_CUDA;
		if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { g_err = "dwt_decode: stream error"; rc = GB200_ERR_CUDA; }
	}
	gb200_plan_destroy(pl);
	return rc;
}

int gb200_t1_encode_blocks(gb200_ctx *ctx, const int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, int rate_control, uint32_t max_passes, gb200_cblk_enc *results, uint32_t *rates,
		double *dists, uint8_t *data, uint64_t data_capacity, uint64_t *data_len) {
	if (!ctx || !plane || !blocks || !results || !rates || !dists || !data_len || !max_passes) FAIL(GB200_ERR_PARAM, "bad arguments");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	std::vector<EncBlock> eb(nblocks);
	uint64_t off = 0, soff = 0;
	bool styles = false;
	DevBuf d_plane, d_blocks, d_results, d_rates, d_dists, d_scratch, d_data, d_symbols;
	auto freeall = [&]() { d_plane.release(); d_blocks.release(); d_results.release(); d_rates.release(); d_dists.release(); d_scratch.release(); d_data.release(); d_symbols.release(); };
	if (d_plane.alloc(std::max<size_t>((size_t) width * height * 4, 16))) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
	for (uint32_t i = 0; i < nblocks; ++i) {
		const gb200_t1_block &b = blocks[i];
		if (b.w > 64 || b.h > 64 || b.x + b.w > width || b.y + b.h > height || b.orient > 3) { freeall(); FAIL(GB200_ERR_PARAM, "bad block"); }
		EncBlock &e = eb[i];
		memset(&e, 0, sizeof(e));
		e.src = (const int32_t*) d_plane.p + (size_t) b.y * width + b.x;
		e.stride = width; e.w = (uint16_t) b.w; e.h = (uint16_t) b.h; e.orient = (uint8_t) b.orient;
		e.reversible = b.qmfbid == 1; e.inv_step = (int32_t) b.inv_step;
		if (b.cblk_sty & ~(uint32_t) STY_ALL) { freeall(); FAIL(GB200_ERR_UNSUPPORTED, "the block-list entry points take Part-1 blocks; HT blocks go through a plan"); }
		e.sty = (uint8_t) b.cblk_sty;
		styles |= b.cblk_sty != 0;
		e.pass_offset = i * max_passes; e.max_passes = max_passes;
		e.scratch_cap = (uint32_t) align_up((uint64_t) b.w * b.h * 4 + 2 + (b.cblk_sty ? 4 * max_passes : 0), 16);
		e.scratch_off = off; e.rd_weight = b.rd_weight;
		e.sym_off = soff;
		e.sym_cap = t1_symbol_capacity(b.w, b.h, (max_passes + 2) / 3);
		soff += e.sym_cap;
		off += e.scratch_cap;
	}
	size_t np = (size_t) nblocks * max_passes;
	if (d_blocks.alloc(std::max<size_t>(nblocks, 1) * sizeof(EncBlock)) || d_results.alloc(std::max<size_t>(nblocks, 1) * sizeof(EncResult))
			|| d_rates.alloc(std::max<size_t>(np, 1) * 4) || d_dists.alloc(std::max<size_t>(np, 1) * 8) || d_scratch.alloc(std::max<uint64_t>(off, 16))
			|| d_data.alloc(std::max<uint64_t>(off, 16)) || d_symbols.alloc(std::max<uint64_t>(soff, 16) + 64)) { freeall(); FAIL(GB200_ERR_NOMEM, "cudaMalloc failed"); }
	cudaMemcpyAsync(d_plane.p, plane, (size_t) width * height * 4, cudaMemcpyHostToDevice, s);
	cudaMemcpyAsync(d_blocks.p, eb.data(), nblocks * sizeof(EncBlock), cudaMemcpyHostToDevice, s);
	cudaMemsetAsync(d_scratch.p, 0, d_scratch.bytes, s);
	cudaMemsetAsync(d_rates.p, 0, d_rates.bytes, s);
	cudaMemsetAsync(d_dists.p, 0, d_dists.bytes, s);
	launch_t1_encode((const EncBlock*) d_blocks.p, nblocks, rate_control, styles ? 1 : 0, (uint8_t*) d_symbols.p, (uint8_t*) d_scratch.p,
			(EncResult*) d_results.p, (uint32_t*) d_rates.p, (double*) d_dists.p, s);
	launch_t1_gather((const EncBlock*) d_blocks.p, (EncResult*) d_results.p, nblocks, (const uint8_t*) d_scratch.p, (uint8_t*) d_data.p, nullptr, s);
	int rc = launch_check(ctx, 4);
	if (!rc && nblocks) {
		cudaMemcpyAsync(results, d_results.p, nblocks * sizeof(EncResult), cudaMemcpyDeviceToHost, s);
		cudaMemcpyAsync(rates, d_rates.p, np * 4, cudaMemcpyDeviceToHost, s);
		cudaMemcpyAsync(dists, d_dists.p, np * 8, cudaMemcpyDeviceToHost, s);
	}
	cudaError_t e = cudaStreamSynchronize(s);
	if (!rc && e != cudaSuccess) { g_err = std::string("t1 encode: ") + cudaGetErrorString(e); rc = GB200_ERR_CUDA; }
	uint64_t total = 0;
	if (!rc) {
		for (uint32_t i = 0; i < nblocks; ++i) {
			if (results[i].numpasses == 0xFFFFFFFFu) { g_err = "a code block overflowed its byte or pass budget"; rc = GB200_ERR_CAPACITY; break; }
			total = std::max<uint64_t>(total, results[i].data_offset + results[i].data_len);
		}
		*data_len = total;
		if (!rc && (total > data_capacity || (total && !data))) { g_err = "data buffer too small"; rc = GB200_ERR_CAPACITY; }
		if (!rc && total) {
			cudaMemcpyAsync(data, d_data.p, total, cudaMemcpyDeviceToHost, s);
			if (cudaStreamSynchronize(s) != cudaSuccess) { g_err = "t1 encode: download failed"; rc = GB200_ERR_CUDA; }
		}
	}
	freeall();
	return rc;
}

int gb200_t1_decode_blocks(gb200_ctx *ctx, int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, const gb200_cblk_dec *inputs, const uint8_t *data, uint64_t data_len) {
	return gb200_t1_decode_blocks_segs(ctx, plane, width, height, nblocks, blocks, inputs, nullptr, nullptr, data, data_len);
}

int gb200_t1_decode_blocks_segs(gb200_ctx *ctx, int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, const gb200_cblk_dec *inputs, const uint32_t *seg_start, const gb200_cblk_seg *segs,
		const uint8_t *data, uint64_t data_len) {
	if (!ctx || !plane || !blocks || !inputs || (seg_start && !segs)) FAIL(GB200_ERR_PARAM, "bad arguments");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	std::vector<DecBlock> db(nblocks);
	DevBuf d_plane, d_blocks, d_inputs, d_data, d_seg_start, d_segs;
	auto freeall = [&]() { d_plane.release(); d_blocks.release(); d_inputs.release(); d_data.release(); d_seg_start.release(); d_segs.release(); };
	bool styles = false;
	if (d_plane.alloc(std::max<size_t>((size_t) width * height * 4, 16))) FAIL(GB200_ERR_NOMEM, "cudaMalloc failed");
	uint32_t maxw = 1, maxh = 1;
	for (uint32_t i = 0; i < nblocks; ++i) {
		const gb200_t1_block &b = blocks[i];
		if (b.w > 64 || b.h > 64 || b.x + b.w > width || b.y + b.h > height || b.orient > 3) { freeall(); FAIL(GB200_ERR_PARAM, "bad block"); }
		if (inputs[i].data_len && inputs[i].data_offset + inputs[i].data_len > data_len) { freeall(); FAIL(GB200_ERR_PARAM, "block bytes outside the data buffer"); }
		if (inputs[i].numbps + b.roishift > 30) { freeall(); FAIL(GB200_ERR_UNSUPPORTED, "more than 30 bit planes (t1.cpp:1056)"); }
		DecBlock &d = db[i];
		memset(&d, 0, sizeof(d));
		d.dst = (int32_t*) d_plane.p + (size_t) b.y * width + b.x;
		d.stride = width; d.w = (uint16_t) b.w; d.h = (uint16_t) b.h; d.orient = (uint8_t) b.orient;
		d.reversible = b.qmfbid == 1; d.stepsize = b.stepsize;
		if (b.cblk_sty & ~(uint32_t) STY_ALL) { freeall(); FAIL(GB200_ERR_UNSUPPORTED, "the block-list entry points take Part-1 blocks; HT blocks go through a plan"); }
		d.sty = (uint8_t) b.cblk_sty;
		styles |= b.cblk_sty != 0;
		if (b.roishift > 29) { freeall(); FAIL(GB200_ERR_UNSUPPORTED, "ROI shift of 30 or more bit planes"); }
		d.roishift = (uint8_t) b.roishift;
		maxw = std::max(maxw, b.w); maxh = std::max(maxh, b.h);
	}
	if (d_blocks.alloc(std::max<size_t>(nblocks, 1) * sizeof(DecBlock)) || d_inputs.alloc(std::max<size_t>(nblocks, 1) * sizeof(DecInput))
			|| d_data.alloc(align_up(data_len + T1_DEC_DATA_SLACK, 256))) { freeall(); FAIL(GB200_ERR_NOMEM, "cudaMalloc failed"); }
	cudaMemcpyAsync(d_plane.p, plane, (size_t) width * height * 4, cudaMemcpyHostToDevice, s);
	cudaMemcpyAsync(d_blocks.p, db.data(), nblocks * sizeof(DecBlock), cudaMemcpyHostToDevice, s);
	cudaMemcpyAsync(d_inputs.p, inputs, nblocks * sizeof(DecInput), cudaMemcpyHostToDevice, s);
	if (data_len) cudaMemcpyAsync(d_data.p, data, data_len, cudaMemcpyHostToDevice, s);
	if (seg_start) {
		const size_t ns = seg_start[nblocks];
		if (d_seg_start.alloc(((size_t) nblocks + 1) * 4) || d_segs.alloc(std::max<size_t>(ns, 1) * sizeof(DecSeg))) { freeall(); FAIL(GB200_ERR_NOMEM, "cudaMalloc failed"); }
		cudaMemcpyAsync(d_seg_start.p, seg_start, ((size_t) nblocks + 1) * 4, cudaMemcpyHostToDevice, s);
		if (ns) cudaMemcpyAsync(d_segs.p, segs, ns * sizeof(DecSeg), cudaMemcpyHostToDevice, s);
	}
	int rc = GB200_OK;
	if (launch_t1_decode((const DecBlock*) d_blocks.p, (const DecInput*) d_inputs.p, nblocks, (const uint8_t*) d_data.p, maxw, maxh,
			styles ? 1 : 0, seg_start ? (const uint32_t*) d_seg_start.p : nullptr, seg_start ? (const DecSeg*) d_segs.p : nullptr, s)) {
		g_err = "Tier-1 decode: code block state does not fit in shared memory";
		rc = GB200_ERR_UNSUPPORTED;
	} else rc = launch_check(ctx, T1_DEC_LAUNCHES);
	cudaMemcpyAsync(plane, d_plane.p, (size_t) width * height * 4, cudaMemcpyDeviceToHost, s);
	cudaError_t e = cudaStreamSynchronize(s);
	if (!rc && e != cudaSuccess) { g_err = std::string("t1 decode: ") + cudaGetErrorString(e); rc = GB200_ERR_CUDA; }
	freeall();
	return rc;
}

} // extern "C"
