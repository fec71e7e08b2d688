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
 *   grk::Tier1::decodeCodeblocks(...)                 Tier1.cpp:177            gb200_t1_decode_blocks
 *   grk::Wavelet::decode(...)                         Wavelet.cpp:47           gb200_dwt_decode
 *   grk::mct::decode_rev / decode_irrev               mct.cpp:143, 352         gb200_mct_decode_*
 *   grk::TileProcessor::dc_level_shift_decode()       TileProcessor.cpp:1377   gb200_dc_shift_decode
 *
 * The encode side keeps a tile resident on the device from the level shift to the code-block bytes
 * (one H2D of the planes, one D2H of bytes + pass tables).  The decode side uses the stage-level entry
 * points on the host tile buffers, one call per reference stage.  GROK_B200_DEVICE selects the GPU.
 */
#include "grok_includes.h"
#include "Tier1.h"
#include "T1Interface.h"
#include "dwt_utils.h"
#include "../include/grok_b200.h"
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace {

gb200_ctx *g_ctx = nullptr;
std::mutex g_mu;
uint64_t g_calls[8] = {0};

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

struct TileResult {
	gb200_plan *plan = nullptr;
	std::vector<gb200_cblk_enc> blocks;
	std::vector<uint32_t> rates;
	std::vector<double> dists;
	std::vector<uint8_t> data;
};
std::map<grk::grk_tcd_tile*, TileResult> g_results;

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

namespace grk {

/* ---- encode ------------------------------------------------------------------------------------ */

bool TileProcessor::dc_level_shift_encode() {
	g_calls[0]++;
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
	TileResult &R = g_results[tile];
	if (R.plan) { gb200_plan_destroy(R.plan); R.plan = nullptr; }
	if (gb200_plan_create(ctx(), 1, &tp, 1, &R.plan) != GB200_OK) fail("gb200_plan_create");
	R.blocks.resize(gb200_plan_num_blocks(R.plan));
	R.rates.resize(gb200_plan_num_pass_slots(R.plan) + 1);
	R.dists.resize(gb200_plan_num_pass_slots(R.plan) + 1);
	R.data.resize(gb200_plan_data_capacity(R.plan) + 16);
	uint64_t len = 0;
	if (gb200_encode_tiles(R.plan, planes.data(), R.blocks.data(), R.rates.data(), R.dists.data(), R.data.data(), R.data.size(), &len) != GB200_OK)
		fail("gb200_encode_tiles");
	return true;
}

bool TileProcessor::mct_encode() { g_calls[1]++; return true; }
bool TileProcessor::dwt_encode() { g_calls[2]++; return true; }

bool Tier1::encodeCodeblocks(grk_tcp *tcp, grk_tcd_tile *tile, const double *, uint32_t, bool doRateControl) {
	g_calls[3]++;
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
						}
						if (doRateControl && e.numpasses) tile->distotile += R.dists[po + e.numpasses - 1];
					}
				}
			}
		}
	}
	gb200_plan_destroy(R.plan);
	g_results.erase(it);
	return true;
}

/* ---- decode ------------------------------------------------------------------------------------ */

bool Tier1::decodeCodeblocks(grk_tcp *, uint16_t, uint16_t, std::vector<decodeBlockInfo*> *blocks) {
	g_calls[4]++;
	if (!blocks || blocks->empty()) return true;
	const size_t n = blocks->size();
	auto tilec = (*blocks)[0]->tilec;
	int32_t *plane = tilec->buf->get_ptr(0, 0, 0, 0);
	const uint32_t width = tilec->width(), height = tilec->height();
	std::vector<gb200_t1_block> desc(n);
	std::vector<gb200_cblk_dec> in(n);
	std::vector<uint8_t> data;
	std::vector<uint32_t> seg_start(n + 1, 0);
	std::vector<gb200_cblk_seg> segs;
	for (size_t i = 0; i < n; ++i) {
		auto b = (*blocks)[i];
		auto cblk = b->cblk;
		gb200_t1_block &d = desc[i];
		memset(&d, 0, sizeof(d));
		d.x = b->x; d.y = b->y; d.w = cblk->x1 - cblk->x0; d.h = cblk->y1 - cblk->y0;
		d.orient = b->bandno; d.qmfbid = b->qmfbid; d.stepsize = b->stepsize;
		d.cblk_sty = b->cblk_sty;
		d.roishift = b->roishift;
		gb200_cblk_dec &c = in[i];
		memset(&c, 0, sizeof(c));
		if (b->cblk_sty & GRK_CBLKSTY_HT) {
			fprintf(stderr, "grok_tcd_shim: HT blocks are outside this build's scope\n");
			abort();
		}
		c.numbps = cblk->numbps - b->roishift;
		c.data_offset = data.size();
		uint32_t passes = 0;
		for (uint32_t s = 0; s < cblk->numSegments; ++s) { /* T1Part1.cpp:160-171 */
			passes += cblk->segs[s].numpasses;
			gb200_cblk_seg sg;
			sg.len = cblk->segs[s].len;
			sg.numpasses = cblk->segs[s].numpasses;
			segs.push_back(sg);
		}
		seg_start[i + 1] = (uint32_t) segs.size();
		c.numpasses = passes;
		for (size_t k = 0; k < cblk->seg_buffers.size(); ++k) { /* T1Part1.cpp:153-158 */
			grk_buf *seg = (grk_buf*) cblk->seg_buffers.get(k);
			data.insert(data.end(), seg->buf, seg->buf + seg->len);
		}
		c.data_len = (uint32_t) (data.size() - c.data_offset);
		delete b;
	}
	if (segs.empty()) segs.resize(1);
	if (gb200_t1_decode_blocks_segs(ctx(), plane, width, height, (uint32_t) n, desc.data(), in.data(), seg_start.data(), segs.data(),
			data.data(), data.size()) != GB200_OK)
		fail("gb200_t1_decode_blocks_segs");
	return true;
}

bool Wavelet::decode(TileProcessor *, TileComponent *tilec, uint32_t numres, uint8_t qmfbid) {
	g_calls[5]++;
	auto full = tilec->resolutions + tilec->numresolutions - 1;
	if (gb200_dwt_decode(ctx(), tilec->buf->get_ptr(0, 0, 0, 0), full->x0, full->y0, full->x1, full->y1, tilec->numresolutions,
			numres, qmfbid) != GB200_OK)
		fail("gb200_dwt_decode");
	return true;
}

void mct::decode_rev(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) {
	g_calls[6]++;
	if (gb200_mct_decode_rev(ctx(), c0, c1, c2, n) != GB200_OK) fail("gb200_mct_decode_rev");
}

void mct::decode_irrev(float *c0, float *c1, float *c2, uint64_t n) {
	g_calls[6]++;
	if (gb200_mct_decode_irrev(ctx(), c0, c1, c2, n) != GB200_OK) fail("gb200_mct_decode_irrev");
}

bool TileProcessor::dc_level_shift_decode() {
	g_calls[7]++;
	for (uint32_t c = 0; c < tile->numcomps; ++c) {
		auto tilec = tile->comps + c;
		auto tccp = m_tcp->tccps + c;
		auto ic = image->comps + c;
		const int32_t lo = ic->sgnd ? -(1 << (ic->prec - 1)) : 0;
		const int32_t hi = ic->sgnd ? (1 << (ic->prec - 1)) - 1 : (int32_t) ((1u << ic->prec) - 1);
		const uint64_t n = (uint64_t) tilec->buf->reduced_image_dim.width() * tilec->buf->reduced_image_dim.height();
		if (gb200_dc_shift_decode(ctx(), tilec->buf->get_ptr(0, 0, 0, 0), n, tccp->m_dc_level_shift, tccp->qmfbid, lo, hi) != GB200_OK)
			fail("gb200_dc_shift_decode");
	}
	return true;
}

} // namespace grk
