/*
 * grok_tcd_shim.cpp -- the reference-side binding of libgrok_b200.so at Grok's TCD stage seam
 * ("B2").  This file is what a Grok maintainer adds: it is compiled against
 * Grok's own private headers and DEFINES the stage functions that TileProcessor::encode_tile /
 * decode_tile call, forwarding them to the C ABI of include/grok_b200.h.  Loaded in front of the stock
 * libgrok (LD_PRELOAD, or dlopen(RTLD_GLOBAL) before libgrok), the dynamic linker resolves the
 * TCD's calls to these definitions, so an UNMODIFIED Grok keeps Tier-2, PCRD and codestream I/O and
 * runs level shift, MCT, DWT, quantisation and Tier-1 on the B200.
 *
 *   replaced symbol                                   reference definition
 *   grk::TileProcessor::dc_level_shift_encode()       TileProcessor.cpp:1449   (whole encode path runs here)
 *   grk::TileProcessor::mct_encode()                  TileProcessor.cpp:1473   (already done -> true)
 *   grk::TileProcessor::dwt_encode()                  TileProcessor.cpp:1520   (already done -> true)
 *   grk::Tier1::encodeCodeblocks(...)                 Tier1.cpp:24             (hands the device results to the host blocks)
 *   grk::Tier1::decodeCodeblocks(...)                 Tier1.cpp:177            (collects the component's blocks and segments)
 *   grk::Wavelet::decode(...)                         Wavelet.cpp:47           (notes the resolutions to reconstruct)
 *   grk::TileProcessor::mct_decode()                  TileProcessor.cpp:1303   gb200_decode_tiles: the whole tile in one batch
 *   grk::TileProcessor::dc_level_shift_decode()       TileProcessor.cpp:1377   (already done -> true)
 *
 * Both directions keep a tile resident on the device for the whole path: encode = one H2D of the planes, one D2H of
 * bytes + pass tables; decode = one H2D of the tile's code-block bytes, one D2H of the finished planes.  The reference
 * walks a tile component by component (TileProcessor.cpp:1141-1177); the shim defers the per-component calls and runs
 * Tier-1, de-quantisation, inverse DWT, inverse MCT and level shift of ALL components of the tile in one batch when the
 * host reaches mct_decode, so that the serial Tier-1 chains of every component overlap.  Plans (geometry, block tables,
 * device buffers) are cached by tile geometry, so equal tiles cost no allocation.  Tiles coded with the HTJ2K block coder
 * (cblk_sty 0x40, grk_compress -M 64) take the same path: the library's HT cleanup-pass kernels stand in for T1HT.  Region (window) decodes
 * (grk_set_decode_area) reconstruct the whole tile on the device and cut the window out.  GROK_B200_DEVICE selects the GPU.
 */
#include "grok_includes.h"
#include "Tier1.h"
#include "T1Interface.h"
#include "dwt_utils.h"
#include "RateControl.h"
#include "../include/grok_b200.h"
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <unordered_set>
#include <mutex>
#include <vector>

namespace {

gb200_ctx *g_ctx = nullptr;
std::mutex g_mu, g_mu2, g_dev_mu; /* g_dev_mu: one tile at a time on the device (cached plans are shared between equal tiles) */
void fail(const char *what);
uint64_t g_calls[8] = {0};
uint64_t g_hulls = 0;
double g_secs[8] = {0}; /* wall clock spent inside the bound stage calls, same indices as g_calls */
struct Timer {
	int i;
	std::chrono::steady_clock::time_point t0;
	explicit Timer(int idx) : i(idx), t0(std::chrono::steady_clock::now()) {}
	~Timer() { g_secs[i] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

gb200_ctx *ctx() {
	std::lock_guard<std::mutex> lk(g_mu);
	if (!g_ctx) {
		const char *d = getenv("GROK_B200_DEVICE");
		if (gb200_create(d ? atoi(d) : 0, &g_ctx) != GB200_OK) {
			fprintf(stderr, "grok_tcd_shim: %s\n", gb200_last_error());
			abort(); /* no CPU fallback inside this layer */
		}
	}
	return g_ctx;
}

/* plans cached by the bytes of their parameters: equal tiles (every interior tile of an image) share one */
struct CachedPlan {
	std::vector<uint8_t> key;
	gb200_plan *plan;
	uint64_t stamp;
};
std::vector<CachedPlan> g_cache[2]; /* [0] decoder, [1] encoder */
uint64_t g_stamp = 0;

gb200_plan *cached_plan(const gb200_tile_params &tp, bool encoder) {
	std::vector<uint8_t> key(sizeof(gb200_tile_params) + tp.numcomps * sizeof(gb200_comp_params));
	gb200_tile_params head = tp;
	head.comps = nullptr;
	memcpy(key.data(), &head, sizeof(head));
	memcpy(key.data() + sizeof(head), tp.comps, tp.numcomps * sizeof(gb200_comp_params));
	std::lock_guard<std::mutex> lk(g_mu2);
	auto &cache = g_cache[encoder ? 1 : 0];
	for (auto &c : cache)
		if (c.key == key) { c.stamp = ++g_stamp; return c.plan; }
	if (cache.size() >= 12) { /* drop the least recently used */
		size_t lru = 0;
		for (size_t i = 1; i < cache.size(); ++i) if (cache[i].stamp < cache[lru].stamp) lru = i;
		gb200_plan_destroy(cache[lru].plan);
		cache.erase(cache.begin() + (long) lru);
	}
	gb200_plan *plan = nullptr;
	if (gb200_plan_create(ctx(), 1, &tp, encoder ? 1 : 0, &plan) != GB200_OK) fail("gb200_plan_create");
	cache.push_back({std::move(key), plan, ++g_stamp});
	return plan;
}

/* what Tier1::decodeCodeblocks saw for one tile component, kept until the tile is complete */
struct PendingBlock {
	uint32_t resno, bandno, x0, y0; /* band coordinates of the block: the key into the plan's table */
	uint32_t numbps, numpasses;
	uint64_t data_offset, data_len;
	uint32_t seg_first, seg_count;
};
struct PendingComp {
	std::vector<PendingBlock> blocks;
	std::vector<gb200_cblk_seg> segs;
	std::vector<uint8_t> data;
	uint32_t numres_decode = 0;
	bool dwt_seen = false;
	/* the band step sizes, taken while tilec->resolutions is alive: TileComponent::release_mem() frees it right after
	 * Wavelet::decode (TileProcessor.cpp:1166) */
	float stepsize[GB200_MAX_BANDS];
	uint32_t band_numbps[GB200_MAX_BANDS]; /* band->numbps: the HT decoder's missing MSBs are counted from it (Tier1.cpp:166) */
};
std::map<grk::TileComponent*, PendingComp> g_pending;
std::map<grk::grk_tcd_tile*, bool> g_tile_done; /* tiles whose level shift already happened on the device */

struct TileResult {
	gb200_plan *plan = nullptr;
	std::vector<gb200_cblk_enc> blocks;
	std::vector<uint32_t> rates;
	std::vector<double> dists;
	std::vector<uint16_t> slopes; /* feasible truncation points from the device (empty: the host computes them) */
	std::vector<uint8_t> data;
};
std::map<grk::grk_tcd_tile*, TileResult> g_results;

/* pass tables whose slopes came from the device (gb200_encode_slopes): RateControl::convexHull below skips those.
 * GROK_B200_HOST_HULL=1 leaves the convex hull to the host (A/B). */
std::mutex g_hull_mu;
std::unordered_set<const grk::grk_tcd_pass*> g_hulled;
bool device_hull() { static const bool on = !getenv("GROK_B200_HOST_HULL"); return on; }

/* which coding passes end a codeword segment: t1_enc_is_term_pass, t1.cpp:1131-1151 (pass 0 is the cleanup pass of the
 * top bit plane, then significance / refinement / cleanup per lower plane) */
bool is_term_pass(uint32_t numbps, uint32_t cblksty, uint32_t passno) {
	const int32_t bpno = (int32_t) numbps - 1 - (int32_t) ((passno + 2) / 3);
	const uint32_t passtype = (passno + 2) % 3;
	if (passtype == 2 && bpno == 0) return true;
	if (cblksty & GRK_CBLKSTY_TERMALL) return true;
	if (cblksty & GRK_CBLKSTY_LAZY) {
		if (bpno == (int32_t) numbps - 4 && passtype == 2) return true;
		if (bpno < (int32_t) numbps - 4 && passtype > 0) return true;
	}
	return false;
}

void fail(const char *what) {
	fprintf(stderr, "grok_tcd_shim: %s: %s\n", what, gb200_last_error());
	abort();
}

} // namespace

extern "C" uint64_t grok_b200_shim_calls(int i) { return g_calls[i & 7]; }
extern "C" uint64_t grok_b200_shim_hulls() { return g_hulls; } /* convexHull calls answered with the device's slopes */
extern "C" double grok_b200_shim_seconds(int i) { return g_secs[i & 7]; }
extern "C" void grok_b200_shim_reset_seconds(void) { for (auto &v : g_secs) v = 0; }

namespace grk {

/* ---- encode ------------------------------------------------------------------------------------ */

bool TileProcessor::dc_level_shift_encode() {
	g_calls[0]++;
	Timer timer(0);
	const uint32_t nc = tile->numcomps;
	std::vector<gb200_comp_params> cp(nc);
	std::vector<const int32_t*> planes(nc);
	const double *mct_norms = nullptr;
	uint32_t mct_numcomps = 0;
	if (m_tcp->mct == 1) { /* TileProcessor::t1_encode, TileProcessor.cpp:1540-1552 */
		mct_numcomps = 3;
		mct_norms = m_tcp->tccps->qmfbid == 0 ? mct::get_norms_irrev() : mct::get_norms_rev();
	} else {
		mct_numcomps = image->numcomps;
		mct_norms = (const double*) m_tcp->mct_norms;
	}
	for (uint32_t c = 0; c < nc; ++c) {
		auto tilec = tile->comps + c;
		auto tccp = m_tcp->tccps + c;
		gb200_comp_params &p = cp[c];
		memset(&p, 0, sizeof(p));
		p.x0 = tilec->x0; p.y0 = tilec->y0; p.x1 = tilec->x1; p.y1 = tilec->y1;
		p.numres = tilec->numresolutions;
		p.cblkw_expn = tccp->cblkw; p.cblkh_expn = tccp->cblkh;
		for (uint32_t r = 0; r < p.numres; ++r) { p.prcw_expn[r] = tccp->prcw[r]; p.prch_expn[r] = tccp->prch[r]; }
		p.qmfbid = tccp->qmfbid;
		p.prec = image->comps[c].prec; p.sgnd = image->comps[c].sgnd;
		p.dc_shift = tccp->m_dc_level_shift;
		p.cblk_sty = tccp->cblk_sty; p.roishift = tccp->roishift;
		const double w1 = (mct_norms && c < mct_numcomps) ? mct_norms[c] : 1.0;
		for (uint32_t r = 0; r < p.numres; ++r) {
			auto res = tilec->resolutions + r;
			for (uint32_t b = 0; b < res->numbands; ++b) {
				auto band = res->bands + b;
				const uint32_t bi = r == 0 ? 0 : 3 * r - 2 + b;
				p.stepsize[bi] = band->stepsize;
				p.inv_step[bi] = band->inv_step;
				p.band_numbps[bi] = band->numbps;
				const uint32_t level = p.numres - 1 - r; /* T1Part1.cpp:114 */
				const double w2 = tccp->qmfbid == 1 ? dwt_utils::getnorm_53(level, band->bandno) : dwt_utils::getnorm_97(level, band->bandno);
				p.rd_weight[bi] = w1 * w2 * (double) band->stepsize; /* t1.cpp:928 */
			}
		}
		planes[c] = tilec->buf->get_ptr(0, 0, 0, 0);
	}
	gb200_tile_params tp;
	memset(&tp, 0, sizeof(tp));
	tp.numcomps = nc;
	tp.mct = m_tcp->mct;
	tp.rate_control = needs_rate_control();
	tp.comps = cp.data();
	std::lock_guard<std::mutex> dev_lock(g_dev_mu);
	TileResult &R = g_results[tile];
	R.plan = cached_plan(tp, true); /* owned by the cache */
	R.blocks.resize(gb200_plan_num_blocks(R.plan));
	R.rates.resize(gb200_plan_num_pass_slots(R.plan) + 1);
	R.dists.resize(gb200_plan_num_pass_slots(R.plan) + 1);
	R.data.resize(gb200_plan_data_capacity(R.plan) + 16);
	uint64_t len = 0;
	if (gb200_encode_tiles(R.plan, planes.data(), R.blocks.data(), R.rates.data(), R.dists.data(), R.data.data(), R.data.size(), &len) != GB200_OK)
		fail("gb200_encode_tiles");
	R.slopes.clear();
	if (tp.rate_control && device_hull()) { /* PCRD preparation while the pass tables are still on the device */
		R.slopes.resize(gb200_plan_num_pass_slots(R.plan) + 1);
		if (gb200_encode_slopes(R.plan, R.slopes.data()) != GB200_OK) fail("gb200_encode_slopes");
	}
	return true;
}

bool TileProcessor::mct_encode() { g_calls[1]++; return true; }
bool TileProcessor::dwt_encode() { g_calls[2]++; return true; }

bool Tier1::encodeCodeblocks(grk_tcp *tcp, grk_tcd_tile *tile, const double *, uint32_t, bool doRateControl) {
	g_calls[3]++;
	Timer timer(3);
	auto it = g_results.find(tile);
	if (it == g_results.end()) { fprintf(stderr, "grok_tcd_shim: no device result for this tile\n"); abort(); }
	TileResult &R = it->second;
	const gb200_cblk_info *info = gb200_plan_blocks(R.plan);
	size_t i = 0;
	tile->distotile = 0;
	for (uint32_t compno = 0; compno < tile->numcomps; ++compno) {
		auto tilec = tile->comps + compno;
		const uint32_t tilec_sty = tcp->tccps[compno].cblk_sty;
		for (uint32_t resno = 0; resno < tilec->numresolutions; ++resno) {
			auto res = tilec->resolutions + resno;
			for (uint32_t bandno = 0; bandno < res->numbands; ++bandno) {
				auto band = res->bands + bandno;
				for (uint32_t precno = 0; precno < res->pw * res->ph; ++precno) {
					auto prc = band->precincts + precno;
					for (uint32_t cblkno = 0; cblkno < prc->cw * prc->ch; ++cblkno, ++i) {
						auto cblk = prc->cblks.enc + cblkno;
						if (i >= R.blocks.size() || info[i].x0 != cblk->x0 || info[i].y0 != cblk->y0 || info[i].x1 != cblk->x1
								|| info[i].y1 != cblk->y1 || info[i].compno != compno || info[i].resno != resno) {
							fprintf(stderr, "grok_tcd_shim: block table mismatch at %zu\n", i);
							abort();
						}
						const gb200_cblk_enc &e = R.blocks[i];
						cblk->numbps = e.numbps;
						cblk->num_passes_encoded = e.numpasses;
						if (e.data_len > cblk->data_size) { fprintf(stderr, "grok_tcd_shim: block bytes exceed the host buffer\n"); abort(); }
						if (e.data_len) memcpy(cblk->data, R.data.data() + e.data_offset, e.data_len);
						const uint32_t po = info[i].pass_offset;
						for (uint32_t p = 0; p < e.numpasses; ++p) {
							auto pass = cblk->passes + p;
							pass->rate = R.rates[po + p];
							pass->len = pass->rate - (p ? R.rates[po + p - 1] : 0);
							pass->distortiondec = R.dists[po + p];
							pass->term = is_term_pass(e.numbps, tilec_sty, p) ? 1 : 0;
							if (!R.slopes.empty()) pass->slope = R.slopes[po + p];
						}
						if (!R.slopes.empty() && e.numpasses) {
							std::lock_guard<std::mutex> lk(g_hull_mu);
							if (g_hulled.size() > (1u << 20)) g_hulled.clear(); /* tables the host never asked about (rate control off later) */
							g_hulled.insert(cblk->passes);
						}
						if (doRateControl && e.numpasses) tile->distotile += R.dists[po + e.numpasses - 1];
					}
				}
			}
		}
	}
	g_results.erase(it);
	return true;
}

/* RateControl::convexHull(passes, n), t2/RateControl.cpp:31, called per block by the rate allocator (TileProcessor.cpp:409):
 * the slopes of a table filled above are already in place; anything else goes to the host's own code. */
void RateControl::convexHull(grk_tcd_pass *pass, uint32_t numPasses) {
	{
		std::lock_guard<std::mutex> lk(g_hull_mu);
		auto it = g_hulled.find(pass);
		if (it != g_hulled.end()) { g_hulled.erase(it); g_hulls++; return; }
	}
	using Fn = void (*)(grk_tcd_pass*, uint32_t);
	static Fn real = (Fn) dlsym(RTLD_NEXT, "_ZN3grk11RateControl10convexHullEPNS_12grk_tcd_passEj");
	if (!real) { fprintf(stderr, "grok_tcd_shim: the host's RateControl::convexHull is not visible\n"); abort(); }
	real(pass, numPasses);
}

/* ---- decode ------------------------------------------------------------------------------------ */

bool Tier1::decodeCodeblocks(grk_tcp *, uint16_t, uint16_t, std::vector<decodeBlockInfo*> *blocks) {
	g_calls[4]++;
	Timer timer(4);
	if (!blocks || blocks->empty()) return true;
	auto tilec = (*blocks)[0]->tilec;
	/* only collect; the tile runs as one batch in TileProcessor::mct_decode */
	std::lock_guard<std::mutex> lk(g_mu2);
	PendingComp &P = g_pending[tilec];
	P.blocks.clear(); P.segs.clear(); P.data.clear();
	P.blocks.reserve(blocks->size());
	for (auto b : *blocks) {
		auto cblk = b->cblk;
		PendingBlock pb;
		pb.resno = b->resno; pb.bandno = b->bandno; pb.x0 = cblk->x0; pb.y0 = cblk->y0;
		pb.numbps = cblk->numbps - b->roishift;
		pb.seg_first = (uint32_t) P.segs.size();
		pb.numpasses = 0;
		for (uint32_t s = 0; s < cblk->numSegments; ++s) {
			gb200_cblk_seg sg;
			sg.len = cblk->segs[s].len;
			sg.numpasses = cblk->segs[s].numpasses;
			pb.numpasses += sg.numpasses;
			P.segs.push_back(sg);
		}
		pb.seg_count = (uint32_t) P.segs.size() - pb.seg_first;
		pb.data_offset = P.data.size();
		for (size_t k = 0; k < cblk->seg_buffers.size(); ++k) {
			grk_buf *seg = (grk_buf*) cblk->seg_buffers.get(k);
			P.data.insert(P.data.end(), seg->buf, seg->buf + seg->len);
		}
		pb.data_len = P.data.size() - pb.data_offset;
		P.blocks.push_back(pb);
		delete b;
	}
	return true;
}

bool Wavelet::decode(TileProcessor *, TileComponent *tilec, uint32_t numres, uint8_t qmfbid) {
	g_calls[5]++;
	{ /* deferred to mct_decode */
		std::lock_guard<std::mutex> lk(g_mu2);
		PendingComp &P = g_pending[tilec];
		P.numres_decode = numres;
		P.dwt_seen = true;
		for (uint32_t r = 0; r < tilec->numresolutions; ++r) {
			auto res = tilec->resolutions + r;
			for (uint32_t b = 0; b < res->numbands; ++b) {
				P.stepsize[r == 0 ? 0 : 3 * r - 2 + b] = res->bands[b].stepsize; /* carries the x0.5 (and the HT scaling, Quantizer.cpp:98-104) */
				P.band_numbps[r == 0 ? 0 : 3 * r - 2 + b] = res->bands[b].numbps;
			}
		}
		return true;
	}
}

bool TileProcessor::mct_decode() {
	g_calls[6]++;
	Timer timer(6);
	const uint32_t nc = tile->numcomps;
	/* ---- the whole tile in one batch: parameters as on the encode side, step sizes as the decoder derived them ---- */
	std::vector<gb200_comp_params> cp(nc);
	std::vector<PendingComp*> pend(nc, nullptr);
	uint32_t numres_decode = 0;
	{
		std::lock_guard<std::mutex> lk(g_mu2);
		for (uint32_t c = 0; c < nc; ++c) {
			auto it = g_pending.find(tile->comps + c);
			if (it != g_pending.end()) pend[c] = &it->second;
		}
	}
	for (uint32_t c = 0; c < nc; ++c) {
		auto tilec = tile->comps + c;
		auto tccp = m_tcp->tccps + c;
		gb200_comp_params &p = cp[c];
		memset(&p, 0, sizeof(p));
		/* on the decode side tilec->x0.. describe the highest DECODED resolution (TileComponent::finalizeCoordinates,
		 * TileComponent.cpp:128-140); the plan wants the full-resolution rectangle */
		p.x0 = (uint32_t) tilec->unreduced_tile_dim.x0; p.y0 = (uint32_t) tilec->unreduced_tile_dim.y0;
		p.x1 = (uint32_t) tilec->unreduced_tile_dim.x1; p.y1 = (uint32_t) tilec->unreduced_tile_dim.y1;
		p.numres = tilec->numresolutions;
		p.cblkw_expn = tccp->cblkw; p.cblkh_expn = tccp->cblkh;
		for (uint32_t r = 0; r < p.numres; ++r) { p.prcw_expn[r] = tccp->prcw[r]; p.prch_expn[r] = tccp->prch[r]; }
		p.qmfbid = tccp->qmfbid;
		p.prec = image->comps[c].prec; p.sgnd = image->comps[c].sgnd;
		p.dc_shift = tccp->m_dc_level_shift;
		p.cblk_sty = tccp->cblk_sty; p.roishift = tccp->roishift;
		if (!pend[c] || !pend[c]->dwt_seen) { fprintf(stderr, "grok_tcd_shim: tile component reached mct_decode without Wavelet::decode\n"); abort(); }
		for (uint32_t bi = 0; bi < 3 * p.numres - 2; ++bi) {
			p.stepsize[bi] = pend[c]->stepsize[bi];
			p.inv_step[bi] = 8192;
			p.band_numbps[bi] = pend[c]->band_numbps[bi];
			p.rd_weight[bi] = 1.0;
		}
		const uint32_t nd = pend[c]->numres_decode;
		if (c == 0) numres_decode = nd;
		else if (nd != numres_decode) { fprintf(stderr, "grok_tcd_shim: components decoded at different resolutions\n"); abort(); }
	}
	gb200_tile_params tp;
	memset(&tp, 0, sizeof(tp));
	tp.numcomps = nc;
	tp.mct = m_tcp->mct;
	tp.numres_decode = numres_decode;
	tp.comps = cp.data();
	if (tp.mct > 1) { fprintf(stderr, "grok_tcd_shim: array based MCT is outside this build's scope\n"); abort(); }
	std::lock_guard<std::mutex> dev_lock(g_dev_mu);
	gb200_plan *plan = cached_plan(tp, false);
	const size_t nb = gb200_plan_num_blocks(plan);
	const gb200_cblk_info *info = gb200_plan_blocks(plan);
	/* the host's blocks, found in the plan's table by (component, resolution, band, band coordinates) */
	std::vector<gb200_cblk_dec> in(nb);
	std::vector<uint32_t> seg_start(nb + 1, 0);
	std::vector<gb200_cblk_seg> segs;
	std::vector<uint8_t> data;
	memset(in.data(), 0, nb * sizeof(gb200_cblk_dec));
	{
		std::map<uint64_t, const PendingBlock*> byKey;
		size_t i = 0;
		for (uint32_t c = 0; c < nc; ++c) {
			byKey.clear();
			if (pend[c])
				for (auto &pb : pend[c]->blocks)
					byKey[((uint64_t) pb.resno << 58) | ((uint64_t) pb.bandno << 56) | ((uint64_t) pb.x0 << 28) | pb.y0] = &pb;
			const uint64_t base = data.size();
			if (pend[c]) data.insert(data.end(), pend[c]->data.begin(), pend[c]->data.end());
			for (; i < nb && info[i].compno == c; ++i) {
				auto f = byKey.find(((uint64_t) info[i].resno << 58) | ((uint64_t) info[i].bandno << 56) | ((uint64_t) info[i].x0 << 28) | info[i].y0);
				if (f != byKey.end()) {
					const PendingBlock &pb = *f->second;
					in[i].numbps = pb.numbps;
					in[i].numpasses = pb.numpasses;
					in[i].data_len = (uint32_t) pb.data_len;
					in[i].data_offset = base + pb.data_offset;
					for (uint32_t k = 0; k < pb.seg_count; ++k) segs.push_back(pend[c]->segs[pb.seg_first + k]);
				}
				seg_start[i + 1] = (uint32_t) segs.size();
			}
		}
		for (; i < nb; ++i) seg_start[i + 1] = (uint32_t) segs.size();
	}
	if (segs.empty()) segs.resize(1);
	/* a region (window) decode asks for a sub-rectangle of the tile (TileBuffer::reduced_image_dim inside reduced_tile_dim,
	 * TileComponent.cpp:557-583): the device reconstructs the whole tile from the blocks the host selected and the
	 * window is cut out afterwards; a whole-tile decode lands in the tile buffers directly */
	std::vector<int32_t*> planes(nc);
	std::vector<std::vector<int32_t>> whole(whole_tile_decoding ? 0 : nc);
	for (uint32_t c = 0; c < nc; ++c) {
		auto buf = tile->comps[c].buf;
		if (whole_tile_decoding) planes[c] = buf->get_ptr(0, 0, 0, 0);
		else {
			whole[c].resize((size_t) buf->reduced_tile_dim.width() * buf->reduced_tile_dim.height());
			planes[c] = whole[c].data();
		}
	}
	if (gb200_decode_set_segments(plan, seg_start.data(), segs.data()) != GB200_OK) fail("gb200_decode_set_segments");
	if (gb200_decode_tiles(plan, in.data(), data.empty() ? nullptr : data.data(), data.size(), planes.data()) != GB200_OK)
		fail("gb200_decode_tiles");
	if (!whole_tile_decoding)
		for (uint32_t c = 0; c < nc; ++c) {
			auto buf = tile->comps[c].buf;
			const int64_t tw = buf->reduced_tile_dim.width(), ww = buf->reduced_image_dim.width(), wh = buf->reduced_image_dim.height();
			const int64_t ox = buf->reduced_image_dim.x0 - buf->reduced_tile_dim.x0, oy = buf->reduced_image_dim.y0 - buf->reduced_tile_dim.y0;
			if (!buf->data || ww <= 0 || wh <= 0) continue;
			if (ox < 0 || oy < 0 || ox + ww > tw || oy + wh > buf->reduced_tile_dim.height()) { fprintf(stderr, "grok_tcd_shim: decode window outside the tile\n"); abort(); }
			for (int64_t y = 0; y < wh; ++y)
				memcpy(buf->data + y * ww, whole[c].data() + (oy + y) * tw + ox, (size_t) ww * sizeof(int32_t));
		}
	{
		std::lock_guard<std::mutex> lk(g_mu2);
		for (uint32_t c = 0; c < nc; ++c) g_pending.erase(tile->comps + c);
		g_tile_done[tile] = true;
	}
	return true;
}

bool TileProcessor::dc_level_shift_decode() {
	g_calls[7]++;
	std::lock_guard<std::mutex> lk(g_mu2);
	auto it = g_tile_done.find(tile);
	if (it == g_tile_done.end()) { fprintf(stderr, "grok_tcd_shim: level shift reached before the tile was decoded\n"); abort(); }
	g_tile_done.erase(it); /* shifted and clamped on the device already */
	return true;
}

} // namespace grk
