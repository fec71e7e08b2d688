/*
 * grok_plugin_b200.cpp -- libgrok_plugin.so: Grok's official minpf plugin ABI ("B1" of SURVEY.md section 8(b))
 * implemented on top of the C ABI of include/grok_b200.h.
 *
 * An unmodified Grok built with the plugin loader (src/lib/jp2/CMakeLists.txt:188-190) finds this library through
 * grk_plugin_load / `grk_compress -g <dir>` (grok.cpp:834-861), resolves the entry points below by name with dlsym on
 * every call (grok.cpp:862-1103) and, for an encode, hands over the whole front end of the tile coder:
 *
 *   plugin_encode(grk_cparameters*, callback)     plugin_interface.h:74, stub src/lib/jp2_plugin/Plugin.cpp:62-67
 *     - the plugin reads parameters->infile itself (PNM here), runs level shift, RCT/ICT, DWT, quantisation and
 *       Tier-1 on the B200 (gb200_encode_tiles) and describes the result as a grk_plugin_tile tree (grok.h:1223-1278)
 *     - it then calls the host's callback ONCE, synchronously; inside it the host creates its codec and runs
 *       grk_encode_with_plugin: TileProcessor::encode_tile skips its own DC shift / MCT / DWT / T1
 *       (TileProcessor.cpp:994) and pcrd_bisect_* pulls every code block through encode_synch_with_plugin
 *       (plugin_bridge.cpp:148-260), then does PCRD, Tier-2 and the codestream itself.
 *
 * Contract details honoured here (all read in plugin_bridge.cpp):
 *   - the tree is indexed [comp][resno][band][precinct][block] in the host's traversal order;
 *   - the host ALIASES compressedData (cblk->data = plugin_cblk->compressedData, owns_data = false), so every
 *     buffer of the tree stays alive until the callback has returned;
 *   - pass rates are reported one less than the reference's final rate: the host computes
 *     min(rate + 1, total) and then drops a trailing 0xFF (plugin_bridge.cpp:236-244);
 *   - distortionDecrease is cumulative (t1.cpp:1262);
 *   - one grk_plugin_tile describes the whole image (j2k.cpp:2069 re-uses the pointer for every tile), so a
 *     multi-tile request cannot be expressed: plugin_encode returns non-zero and the host falls back to its CPU
 *     path (grk_compress.cpp:2215-2219).  Multi-tile images go through the TCD stage seam instead
 *     (integration/grok_tcd_shim.cpp).
 *
 * Quantisation constants are derived with the host's own code (param_qcd::generate / pull as j2k.cpp:1839-2049
 * does, the three formulas of Quantizer.cpp:65-105, dwt_utils::getnorm_*, mct::get_norms_*): this file is built
 * against Grok's headers and links libgrok, it restates none of its tables.
 * There is no CPU fallback in this layer: plugin_init fails without a CUDA device, and the host then keeps its own path.
 */
#include "grok_includes.h"
#include "dwt_utils.h"
#include "plugin/plugin_interface.h"
#include "../include/grok_b200.h"
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <dirent.h>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

/* one context per device in use: plugin_init(deviceId = -1) takes every GPU of the box (grk_compress -G -1, "A value of -1 will
 * specify all devices", grk_compress.cpp:423-426); g_ctx = the first one, used by the single-image entry points */
std::vector<gb200_ctx*> g_ctxs;
gb200_ctx *g_ctx = nullptr;
bool g_verbose = false;
std::atomic<uint64_t> g_encodes{0}, g_blocks{0}, g_decodes{0};

void say(const char *what) {
	if (g_verbose) fprintf(stderr, "grok_plugin_b200: %s: %s\n", what, gb200_last_error());
}

/* ---- PNM (P5 / P6, 8 or 16 bit big endian): the only input format this adapter reads itself -------------- */
bool pnm_token(FILE *f, uint32_t *v) {
	int c = fgetc(f);
	for (;;) {
		while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(f);
		if (c != '#') break;
		while (c != '\n' && c != EOF) c = fgetc(f);
	}
	if (c < '0' || c > '9') return false;
	uint64_t x = 0;
	while (c >= '0' && c <= '9') { x = x * 10 + (uint64_t) (c - '0'); if (x > 0xFFFFFFFFull) return false; c = fgetc(f); }
	*v = (uint32_t) x; /* the single white-space byte after the token has been consumed */
	return true;
}

grk_image *read_pnm(const grk_cparameters *p) {
	FILE *f = fopen(p->infile, "rb");
	if (!f) return nullptr;
	grk_image *img = nullptr;
	int m0 = fgetc(f), m1 = fgetc(f);
	uint32_t w = 0, h = 0, maxval = 0;
	if (m0 == 'P' && (m1 == '5' || m1 == '6') && pnm_token(f, &w) && pnm_token(f, &h) && pnm_token(f, &maxval) && w && h
			&& maxval && maxval < 65536) {
		const uint32_t nc = m1 == '6' ? 3 : 1;
		uint32_t prec = 1;
		while ((1u << prec) <= maxval) prec++;
		const uint32_t dx = p->subsampling_dx ? p->subsampling_dx : 1, dy = p->subsampling_dy ? p->subsampling_dy : 1;
		std::vector<grk_image_cmptparm> cp(nc);
		for (auto &c : cp) {
			memset(&c, 0, sizeof(c));
			c.dx = dx; c.dy = dy; c.w = w; c.h = h; c.prec = prec; c.sgnd = 0;
			c.x0 = p->image_offset_x0; c.y0 = p->image_offset_y0;
		}
		img = grk_image_create(nc, cp.data(), nc == 3 ? GRK_CLRSPC_SRGB : GRK_CLRSPC_GRAY);
		if (img) {
			/* image area on the reference grid, as the host's PNM reader sets it */
			img->x0 = p->image_offset_x0; img->y0 = p->image_offset_y0;
			img->x1 = img->x0 + (w - 1) * dx + 1; img->y1 = img->y0 + (h - 1) * dy + 1;
			const size_t bps = maxval > 255 ? 2 : 1, row = (size_t) w * nc * bps;
			std::vector<uint8_t> line(row);
			bool ok = true;
			for (uint32_t y = 0; y < h && ok; ++y) {
				ok = fread(line.data(), 1, row, f) == row;
				for (uint32_t x = 0; x < w && ok; ++x)
					for (uint32_t c = 0; c < nc; ++c) {
						const uint8_t *s = line.data() + ((size_t) x * nc + c) * bps;
						img->comps[c].data[(size_t) y * w + x] = bps == 2 ? (s[0] << 8 | s[1]) : s[0];
					}
			}
			if (!ok) { grk_image_destroy(img); img = nullptr; }
		}
	}
	fclose(f);
	return img;
}

/* ---- coding parameters of the single tile, derived the way j2k_setup_encoder + TileComponent::init do ------ */
bool fill_comp(const grk_cparameters *p, const grk_image *img, uint32_t c, bool mct, gb200_comp_params &cp) {
	using namespace grk;
	memset(&cp, 0, sizeof(cp));
	const grk_image_comp *ic = img->comps + c;
	cp.x0 = ceildiv<uint32_t>(img->x0, ic->dx); cp.y0 = ceildiv<uint32_t>(img->y0, ic->dy);
	cp.x1 = ceildiv<uint32_t>(img->x1, ic->dx); cp.y1 = ceildiv<uint32_t>(img->y1, ic->dy);
	cp.numres = p->numresolution;
	cp.cblkw_expn = uint_floorlog2(p->cblockw_init); cp.cblkh_expn = uint_floorlog2(p->cblockh_init);
	const bool rev = !p->irreversible;
	cp.qmfbid = rev ? 1 : 0;
	cp.prec = ic->prec; cp.sgnd = ic->sgnd;
	cp.dc_shift = ic->sgnd ? 0 : 1 << (ic->prec - 1); /* j2k.cpp:1972-1977 */
	cp.cblk_sty = p->cblk_sty;
	cp.roishift = ((int32_t) c == p->roi_compno) ? p->roi_shift : 0;
	/* precinct sizes, j2k.cpp:2001-2048 */
	if ((p->csty & J2K_CCP_CSTY_PRT) && p->res_spec) {
		uint32_t k = 0;
		for (int32_t r = (int32_t) cp.numres - 1; r >= 0; --r, ++k) {
			uint32_t pw, ph;
			if (k < p->res_spec) { pw = p->prcw_init[k]; ph = p->prch_init[k]; }
			else { pw = p->prcw_init[p->res_spec - 1] >> (k - (p->res_spec - 1)); ph = p->prch_init[p->res_spec - 1] >> (k - (p->res_spec - 1)); }
			cp.prcw_expn[r] = pw < 1 ? 1 : uint_floorlog2(pw);
			cp.prch_expn[r] = ph < 1 ? 1 : uint_floorlog2(ph);
		}
	} else
		for (uint32_t r = 0; r < cp.numres; ++r) cp.prcw_expn[r] = cp.prch_expn[r] = 15;
	/* quantisation: QCD generated from component 0 (j2k.cpp:1839-1841), pulled into every component (:2049) */
	const uint8_t numgbits = 2;
	param_qcd qcd;
	qcd.generate(numgbits, cp.numres - 1, rev, img->comps[0].prec, mct, img->comps[0].sgnd);
	grk_stepsize steps[GRK_J2K_MAXBANDS];
	memset(steps, 0, sizeof(steps));
	qcd.pull(steps, rev);
	const double *norms = mct ? (rev ? mct::get_norms_rev() : mct::get_norms_irrev()) : nullptr;
	const double w1 = (norms && c < 3) ? norms[c] : 1.0;
	for (uint32_t r = 0; r < cp.numres; ++r)
		for (uint32_t b = 0; b < (r ? 3u : 1u); ++b) {
			const uint32_t bi = r ? 3 * r - 2 + b : 0, orient = r ? b + 1 : 0;
			const uint32_t gain = rev ? (orient == 0 ? 0 : orient < 3 ? 1 : 2) : 0;
			const uint32_t numbps = ic->prec + gain;
			/* Quantizer::setBandStepSizeAndBps, Quantizer.cpp:65-105 */
			const float stepsize = (float) ((1.0 + steps[bi].mant / 2048.0) * pow(2.0, (int32_t) (numbps - steps[bi].expn)));
			cp.stepsize[bi] = stepsize;
			cp.band_numbps[bi] = cp.roishift + (uint32_t) (steps[bi].expn + numgbits) - 1;
			cp.inv_step[bi] = (uint32_t) ((8192.0 / stepsize) + 0.5f);
			const uint32_t level = cp.numres - 1 - r; /* T1Part1.cpp:114 */
			const double w2 = rev ? dwt_utils::getnorm_53(level, (uint8_t) orient) : dwt_utils::getnorm_97(level, (uint8_t) orient);
			cp.rd_weight[bi] = w1 * w2 * (double) stepsize; /* t1.cpp:912-932 */
		}
	return true;
}

/* plans cached per context by the bytes of their parameters: the frames of a batch share one geometry, and a plan owns
 * gigabytes of device buffers whose allocation costs more than coding a frame */
struct PlanCache {
	struct Entry { std::vector<uint8_t> key; gb200_plan *plan; uint64_t stamp; };
	std::mutex mu;
	std::vector<Entry> entries;
	uint64_t stamp = 0;
	gb200_plan *get(gb200_ctx *ctx, const gb200_tile_params &tp, bool encoder) {
		std::vector<uint8_t> key(sizeof(gb200_tile_params) + tp.numcomps * sizeof(gb200_comp_params) + 1);
		gb200_tile_params head = tp;
		head.comps = nullptr;
		memcpy(key.data(), &head, sizeof(head));
		memcpy(key.data() + sizeof(head), tp.comps, tp.numcomps * sizeof(gb200_comp_params));
		key.back() = encoder ? 1 : 0;
		std::lock_guard<std::mutex> lk(mu);
		for (auto &e : entries)
			if (e.key == key) { e.stamp = ++stamp; return e.plan; }
		if (entries.size() >= 4) {
			size_t lru = 0;
			for (size_t i = 1; i < entries.size(); ++i) if (entries[i].stamp < entries[lru].stamp) lru = i;
			gb200_plan_destroy(entries[lru].plan);
			entries.erase(entries.begin() + (long) lru);
		}
		gb200_plan *plan = nullptr;
		if (gb200_plan_create(ctx, 1, &tp, encoder ? 1 : 0, &plan) != GB200_OK) return nullptr;
		entries.push_back({std::move(key), plan, ++stamp});
		return plan;
	}
	void clear() {
		std::lock_guard<std::mutex> lk(mu);
		for (auto &e : entries) gb200_plan_destroy(e.plan);
		entries.clear();
	}
};
std::vector<std::unique_ptr<PlanCache>> g_caches; /* parallel to g_ctxs */
PlanCache *cache_of(gb200_ctx *ctx) {
	for (size_t i = 0; i < g_ctxs.size(); ++i) if (g_ctxs[i] == ctx) return g_caches[i].get();
	return nullptr;
}

/* the grk_plugin_tile tree over the encoder results; owns every node and the compressed bytes */
struct Tree {
	grk_plugin_tile tile;
	std::vector<grk_plugin_tile_component> comps;
	std::vector<grk_plugin_tile_component*> comp_ptrs;
	std::vector<std::unique_ptr<grk_plugin_resolution[]>> res;
	std::vector<std::vector<grk_plugin_resolution*>> res_ptrs;
	std::vector<std::unique_ptr<grk_plugin_band[]>> bands;
	std::vector<std::vector<grk_plugin_band*>> band_ptrs;
	std::vector<std::unique_ptr<grk_plugin_precinct[]>> precs;
	std::vector<std::vector<grk_plugin_precinct*>> prec_ptrs;
	std::vector<std::vector<grk_plugin_code_block*>> blk_ptrs;
	std::vector<grk_plugin_code_block> blocks;
	std::vector<uint8_t> data;
};

/* the tree, sized from the block table of a plan (host order: comp, resno, band, precinct, block) */
void build_tree(Tree &T, const gb200_cblk_info *info, size_t nb, uint32_t nc, const gb200_comp_params *cps) {
	T.blocks.resize(nb);
	for (auto &b : T.blocks) memset(&b, 0, sizeof(b));
	T.comps.resize(nc);
	T.comp_ptrs.resize(nc);
	for (uint32_t c = 0; c < nc; ++c) {
		const uint32_t numres = cps[c].numres;
		T.res.emplace_back(new grk_plugin_resolution[numres]);
		T.res_ptrs.emplace_back(numres);
		grk_plugin_resolution *res = T.res.back().get();
		for (uint32_t r = 0; r < numres; ++r) {
			const uint32_t nbands = r ? 3 : 1;
			T.bands.emplace_back(new grk_plugin_band[nbands]);
			T.band_ptrs.emplace_back(nbands);
			grk_plugin_band *bands = T.bands.back().get();
			for (uint32_t b = 0; b < nbands; ++b) {
				const uint32_t orient = r ? b + 1 : 0, bi = r ? 3 * r - 2 + b : 0;
				/* every precinct of the resolution's grid exists in the host's tree, also those (and those bands) without a
				 * block: the host indexes precincts[precno] for precno < pw * ph (plugin_bridge.cpp:38-43) */
				uint32_t gpw = 0, gph = 0;
				gb200_precinct_grid(&cps[c], r, &gpw, &gph);
				size_t nprec = (size_t) gpw * gph;
				for (size_t i = 0; i < nb; ++i)
					if (info[i].compno == c && info[i].resno == r && info[i].bandno == orient) nprec = std::max<size_t>(nprec, info[i].precno + 1);
				T.precs.emplace_back(new grk_plugin_precinct[nprec ? nprec : 1]);
				T.prec_ptrs.emplace_back(nprec ? nprec : 1, nullptr);
				grk_plugin_precinct *precs = T.precs.back().get();
				std::vector<size_t> nblk(nprec, 0);
				for (size_t i = 0; i < nb; ++i)
					if (info[i].compno == c && info[i].resno == r && info[i].bandno == orient)
						nblk[info[i].precno] = std::max<size_t>(nblk[info[i].precno], info[i].cblkno + 1);
				for (size_t pr = 0; pr < nprec; ++pr) {
					T.blk_ptrs.emplace_back(nblk[pr] ? nblk[pr] : 1, nullptr); /* never a NULL array */
					precs[pr].numBlocks = nblk[pr];
					precs[pr].blocks = T.blk_ptrs.back().data();
					T.prec_ptrs.back()[pr] = precs + pr;
				}
				for (size_t i = 0; i < nb; ++i)
					if (info[i].compno == c && info[i].resno == r && info[i].bandno == orient) precs[info[i].precno].blocks[info[i].cblkno] = &T.blocks[i];
				bands[b].orient = orient;
				bands[b].numPrecincts = nprec;
				bands[b].precincts = T.prec_ptrs.back().data();
				bands[b].stepsize = cps[c].stepsize[bi];
				T.band_ptrs.back()[b] = bands + b;
			}
			res[r].level = r;
			res[r].numBands = nbands;
			res[r].bands = T.band_ptrs.back().data();
			T.res_ptrs.back()[r] = res + r;
		}
		T.comps[c].numResolutions = numres;
		T.comps[c].resolutions = T.res_ptrs.back().data();
		T.comp_ptrs[c] = &T.comps[c];
	}
	for (size_t i = 0; i < nb; ++i) {
		grk_plugin_code_block &B = T.blocks[i];
		B.x0 = info[i].x0; B.y0 = info[i].y0; B.x1 = info[i].x1; B.y1 = info[i].y1;
		B.numPix = (size_t) (B.x1 - B.x0) * (B.y1 - B.y0);
		B.sortedIndex = (unsigned int) i;
	}
	T.tile.decode_flags = 0;
	T.tile.numComponents = nc;
	T.tile.tileComponents = T.comp_ptrs.data();
}


} // namespace

/* ---- minpf registration (minpf_plugin.h:24-57, loader minpf_plugin_manager.cpp:146-162) --------------------- */
extern "C" PLUGIN_API int32_t grok_b200_plugin_exit() {
	for (auto &c : g_caches) c->clear();
	g_caches.clear();
	for (gb200_ctx *c : g_ctxs) gb200_destroy(c);
	g_ctxs.clear();
	g_ctx = nullptr;
	return 0;
}
static void *plugin_create(grk::minpf_object_params *) { return nullptr; }
static int32_t plugin_destroy(void *) { return 0; }

extern "C" PLUGIN_API grk::minpf_exit_func minpf_post_load_plugin(const char *, const grk::minpf_platform_services *services) {
	grk::minpf_register_params rp;
	rp.version.major = 1; /* must equal the host's, minpf_plugin_manager.cpp:59-61 */
	rp.version.minor = 0;
	rp.createFunc = plugin_create;
	rp.destroyFunc = plugin_destroy;
	if (!services || services->registerObject("GrokB200", &rp) < 0) return nullptr;
	/* Pin this library (and libgrok_b200.so with the CUDA runtime linked into it) for the life of the process: the host
	 * dlclose()s the plugin in grk_plugin_cleanup (minpf_plugin_manager.cpp), and a CUDA runtime that is unloaded and loaded
	 * again while worker threads of the first instance have used it does not survive that (seen as a SIGSEGV in the batch
	 * decode worker of a second load / cleanup cycle).  RTLD_NODELETE makes the later dlclose a reference drop only. */
	Dl_info self;
	if (dladdr((void*) &grok_b200_plugin_exit, &self) && self.dli_fname) dlopen(self.dli_fname, RTLD_NOW | RTLD_NODELETE);
	return grok_b200_plugin_exit;
}

extern "C" PLUGIN_API bool plugin_init(grk_plugin_init_info info) {
	g_verbose = info.verbose;
	if (g_ctx) return true;
	const char *d = getenv("GROK_B200_DEVICE"); /* overrides -G (test harnesses hard-wire deviceId) */
	const int want = d ? atoi(d) : info.deviceId;
	std::vector<int> devices;
	if (want < 0) { /* all devices: the batch entry points deal frames to them round robin */
		const int n = gb200_device_count();
		for (int i = 0; i < n; ++i) devices.push_back(i);
	} else devices.push_back(want);
	for (int dev : devices) {
		gb200_ctx *c = nullptr;
		if (gb200_create(dev, &c) != GB200_OK) { say("plugin_init"); continue; }
		g_ctxs.push_back(c);
		g_caches.emplace_back(new PlanCache());
	}
	if (g_ctxs.empty()) return false; /* false => the host silently keeps its CPU path */
	g_ctx = g_ctxs[0];
	return true;
}

/* counters for the parity tests: 0 = images encoded on the device, 1 = code blocks handed to the host */
/* 3 = contexts (devices) in use, 4 + k = frames coded on device k */
std::atomic<uint64_t> g_per_device[16];
extern "C" PLUGIN_API uint64_t grok_b200_plugin_stat(int i) {
	if (i == 3) return g_ctxs.size();
	if (i >= 4 && i < 20) return g_per_device[i - 4];
	return i == 0 ? g_encodes : i == 1 ? g_blocks : g_decodes;
}

/* ---- encode -------------------------------------------------------------------------------------------------- */
namespace {

/* one image after the device path: the image, the tree over the result buffers, and what the callback needs */
struct EncodedFrame {
	grk_image *img = nullptr;
	Tree T;
	std::string infile, outfile;
	~EncodedFrame() { if (img) grk_image_destroy(img); }
};

/* read p->infile, run DC shift .. Tier-1 on the device, describe the result as a grk_plugin_tile.  0 = ok; 1 = a request
 * this ABI / build cannot express (the host keeps its CPU path); >1 = failure */
int encode_frame(gb200_ctx *ctx, const grk_cparameters *p, EncodedFrame &F) {
	/* what the single grk_plugin_tile of this ABI, or this build of the kernels, cannot express -> host CPU path */
	/* (terminating styles would need pass->term, which encode_synch_with_plugin never sets: plugin_bridge.cpp:148-260) */
	if (p->isHT || p->cblk_sty != 0 || p->decod_format != GRK_PXM_FMT) return 1;
	grk_image *img = read_pnm(p);
	if (!img) return 2;
	F.img = img;
	F.infile = p->infile;
	F.outfile = p->outfile;
	if (p->tile_size_on && (p->cp_tx0 + p->cp_tdx < img->x1 || p->cp_ty0 + p->cp_tdy < img->y1)) return 1; /* more than one tile */
	const uint32_t nc = img->numcomps;
	/* MCT decision of grk_compress.cpp:1996-1998 and j2k.cpp:1961-1970 */
	bool mct = p->tcp_mct == 255 ? nc >= 3 : p->tcp_mct == 1;
	if (p->tcp_mct > 1 && p->tcp_mct != 255) return 1; /* array based MCT */
	if (mct && nc < 3) return 1;
	std::vector<gb200_comp_params> cps(nc);
	for (uint32_t c = 0; c < nc; ++c) fill_comp(p, img, c, mct, cps[c]);
	gb200_tile_params tp;
	memset(&tp, 0, sizeof(tp));
	tp.numcomps = nc;
	tp.mct = mct ? 1 : 0;
	tp.rate_control = 1; /* the host decides later whether it uses the distortions (needs_rate_control) */
	tp.comps = cps.data();
	gb200_plan *plan = cache_of(ctx)->get(ctx, tp, true); /* owned by the context's cache */
	if (!plan) { say("gb200_plan_create"); return 3; }
	const size_t nb = gb200_plan_num_blocks(plan);
	std::vector<gb200_cblk_enc> enc(nb);
	std::vector<uint32_t> rates(gb200_plan_num_pass_slots(plan) + 1);
	std::vector<double> dists(gb200_plan_num_pass_slots(plan) + 1);
	Tree &T = F.T;
	T.data.resize(gb200_plan_data_capacity(plan) + 16);
	std::vector<const int32_t*> planes(nc);
	for (uint32_t c = 0; c < nc; ++c) planes[c] = img->comps[c].data;
	uint64_t len = 0;
	if (gb200_encode_tiles(plan, planes.data(), enc.data(), rates.data(), dists.data(), T.data.data(), T.data.size(), &len) != GB200_OK) {
		say("gb200_encode_tiles");
		return 4;
	}
	const gb200_cblk_info *info = gb200_plan_blocks(plan);
	build_tree(T, info, nb, nc, cps.data());
	for (size_t i = 0; i < nb; ++i) {
		grk_plugin_code_block &B = T.blocks[i];
		const gb200_cblk_enc &e = enc[i];
		B.compressedData = T.data.data() + e.data_offset;
		B.compressedDataLength = e.data_len;
		B.numBitPlanes = e.numbps;
		B.numPasses = e.numpasses;
		if (e.numpasses > 67) return 5;
		const uint32_t po = info[i].pass_offset;
		for (uint32_t k = 0; k < e.numpasses; ++k) {
			const uint32_t r = rates[po + k];
			B.passes[k].distortionDecrease = dists[po + k];
			B.passes[k].rate = r ? r - 1 : 0; /* host: min(rate + 1, total), plugin_bridge.cpp:236 */
			B.passes[k].length = r - (k ? rates[po + k - 1] : 0);
		}
	}
	g_encodes++;
	g_blocks += nb;
	for (size_t k = 0; k < g_ctxs.size() && k < 16; ++k) if (g_ctxs[k] == ctx) g_per_device[k]++;
	return 0;
}

/* the host's turn: header + PCRD + Tier-2 + file write happen inside the callback, synchronously */
int hand_over(EncodedFrame &F, grk_cparameters *p, grk::PLUGIN_ENCODE_USER_CALLBACK callback, bool relative) {
	grk::plugin_encode_user_callback_info cbinfo;
	memset(&cbinfo, 0, sizeof(cbinfo));
	cbinfo.input_file_name = F.infile.c_str();
	cbinfo.outputFileNameIsRelative = relative;
	cbinfo.output_file_name = F.outfile.c_str();
	cbinfo.encoder_parameters = p;
	cbinfo.image = F.img;
	cbinfo.tile = &F.T.tile;
	cbinfo.error_code = 0;
	try {
		callback(&cbinfo);
	} catch (...) {
		return 6; /* nothing in libgrok catches what its callback throws */
	}
	return cbinfo.error_code;
}

} // namespace

extern "C" PLUGIN_API int32_t plugin_encode(grk_cparameters *p, grk::PLUGIN_ENCODE_USER_CALLBACK callback) {
	if (!g_ctx || !p || !callback) return -1;
	EncodedFrame F;
	const int rc = encode_frame(g_ctx, p, F);
	if (rc) return rc;
	return hand_over(F, p, callback, false);
}

/* ---- batch encode: the plugin owns the frame loop (plugin_interface.h:80-86, grk_compress.cpp:2222-2240) -----------------
 * plugin_batch_encode returns at once; one producer thread per device in use (plugin_init with deviceId -1: every GPU of the
 * box) runs the PNM files of input_dir through the device path, frame i on device i mod N, no data crosses between devices; a
 * consumer thread hands the finished frames to the host's callback in file-name order (the callback is the host's
 * single-threaded PCRD / Tier-2 / file writer), so the devices work on the next frames while the host finishes frame i.
 * Frames the ABI cannot express are handed over with tile = NULL: the host's callback then encodes them itself. */
namespace {

struct Batch {
	std::vector<std::thread> producers; /* one per context (device): frame i is coded on device i mod N */
	std::thread consumer;
	std::mutex mu;
	std::condition_variable cv;
	std::vector<std::unique_ptr<EncodedFrame>> ready; /* slot i = frame i once coded */
	size_t next = 0;                                  /* next frame the host gets */
	std::vector<std::string> files;
	grk_cparameters params;
	grk::PLUGIN_ENCODE_USER_CALLBACK callback = nullptr;
	std::atomic<bool> stop{false}, complete{true};
};
Batch *g_batch = nullptr;

bool has_pnm_extension(const std::string &n) {
	const size_t d = n.rfind('.');
	if (d == std::string::npos) return false;
	std::string e = n.substr(d + 1);
	for (auto &c : e) c = (char) tolower(c);
	return e == "pgm" || e == "ppm" || e == "pnm";
}

void batch_join(Batch *b) {
	for (auto &t : b->producers) if (t.joinable()) t.join();
	if (b->consumer.joinable()) b->consumer.join();
}

} // namespace

extern "C" PLUGIN_API int32_t plugin_batch_encode(const char *input_dir, const char *output_dir, grk_cparameters *p,
		grk::PLUGIN_ENCODE_USER_CALLBACK callback) {
	(void) output_dir; /* the host's callback composes the output path from the relative name (grk_compress.cpp:1783-1797) */
	if (!g_ctx || !input_dir || !p || !callback) return -1;
	if (g_batch) {
		if (!g_batch->complete) return -1; /* one batch at a time */
		batch_join(g_batch);
		delete g_batch;
		g_batch = nullptr;
	}
	DIR *d = opendir(input_dir);
	if (!d) return 2;
	Batch *b = new Batch();
	while (dirent *e = readdir(d))
		if (has_pnm_extension(e->d_name)) b->files.push_back(std::string(input_dir) + "/" + e->d_name);
	closedir(d);
	std::sort(b->files.begin(), b->files.end());
	b->params = *p;
	b->callback = callback;
	b->complete = false;
	g_batch = b;
	b->ready.resize(b->files.size());
	const size_t ndev = g_ctxs.size();
	for (size_t k = 0; k < ndev; ++k)
		b->producers.emplace_back([b, k, ndev]() {
			gb200_ctx *ctx = g_ctxs[k];
			for (size_t i = k; i < b->files.size(); i += ndev) {
				{ /* at most two frames per device ahead of the host */
					std::unique_lock<std::mutex> lk(b->mu);
					b->cv.wait(lk, [&]() { return i < b->next + 2 * ndev || b->stop; });
				}
				if (b->stop) break;
				const std::string &f = b->files[i];
				std::unique_ptr<EncodedFrame> F(new EncodedFrame());
				grk_cparameters fp = b->params;
				snprintf(fp.infile, sizeof(fp.infile), "%s", f.c_str());
				const int rc = encode_frame(ctx, &fp, *F);
				F->infile = f;
				F->outfile = f; /* relative: the host keeps the base name */
				if (rc) { F->T.tile.tileComponents = nullptr; F->T.tile.numComponents = 0; if (F->img) { grk_image_destroy(F->img); F->img = nullptr; } }
				std::lock_guard<std::mutex> lk(b->mu);
				b->ready[i] = std::move(F);
				b->cv.notify_all();
			}
		});
	b->consumer = std::thread([b]() {
		for (; b->next < b->files.size();) { /* file-name order, whatever device finished first */
			std::unique_ptr<EncodedFrame> F;
			{
				std::unique_lock<std::mutex> lk(b->mu);
				b->cv.wait(lk, [b]() { return b->ready[b->next] != nullptr || b->stop; });
				if (!b->ready[b->next]) break;
				F = std::move(b->ready[b->next]);
				b->next++;
				b->cv.notify_all();
			}
			grk_cparameters fp = b->params; /* the callback may settle tcp_mct etc. in its copy */
			snprintf(fp.infile, sizeof(fp.infile), "%s", F->infile.c_str());
			grk::plugin_encode_user_callback_info cbinfo;
			memset(&cbinfo, 0, sizeof(cbinfo));
			cbinfo.input_file_name = F->infile.c_str();
			cbinfo.outputFileNameIsRelative = true;
			cbinfo.output_file_name = F->outfile.c_str();
			cbinfo.encoder_parameters = &fp;
			cbinfo.image = F->img; /* NULL + tile NULL: the host loads and encodes the file itself */
			cbinfo.tile = F->img ? &F->T.tile : nullptr;
			try { b->callback(&cbinfo); } catch (...) {}
		}
		b->complete = true;
	});
	return 0;
}

bool decode_batch_complete();
extern "C" PLUGIN_API bool plugin_is_batch_complete(void) { return (!g_batch || g_batch->complete) && decode_batch_complete(); }

extern "C" PLUGIN_API void plugin_stop_batch_encode(void) {
	if (!g_batch) return;
	g_batch->stop = true;
	{ std::lock_guard<std::mutex> lk(g_batch->mu); g_batch->cv.notify_all(); }
	batch_join(g_batch);
	delete g_batch;
	g_batch = nullptr;
}

/* ---- decode --------------------------------------------------------------------------------------------------
 * The staged protocol of grk_decompress.cpp:1336-1560 (decode_callback / pre_decode / post_decode):
 *   1. callback(HEADER) with init_decoders_func set: the host opens its stream, reads the header and calls back
 *      into decode_init below with the header info and the image -> geometry, tree, byte arena;
 *   2. callback(T2) with the tree: the host parses packets and copies every block's segment bytes, bit-plane and
 *      pass counts and the band step sizes into the tree (decode_synch_plugin_with_host, plugin_bridge.cpp:24-87),
 *      then destroys its codec;
 *   3. Tier-1 + de-quantisation + inverse DWT + inverse MCT + level shift/clamp run on the device
 *      (gb200_decode_tiles) straight into image->comps[].data;
 *   4. callback(POST_T1): the host stores the image; callback(CLEAN): it frees it. */
namespace {

struct DecodeJob {
	uint32_t reduce = 0;
	uint32_t nc = 0;
	std::vector<gb200_comp_params> cps;
	gb200_tile_params tp;
	Tree T;
	std::vector<gb200_cblk_info> table; /* every block of the image, host order */
	std::vector<uint64_t> offset; /* of each block in T.data */
	std::vector<uint64_t> capacity; /* bytes reserved for each block */
	uint64_t file_bytes = 0;       /* size of the codestream file: no block carries more */
	bool ready = false;
	int status = 0;
};
thread_local DecodeJob *g_job = nullptr; /* the host calls decode_init back synchronously, on the thread that called its callback */

/* block table of a whole single-tile image (every resolution), component by component: pure geometry, no device */
void enumerate_image(const std::vector<gb200_comp_params> &cps, std::vector<gb200_cblk_info> &out) {
	out.clear();
	for (uint32_t c = 0; c < cps.size(); ++c) {
		const uint64_t n = gb200_enumerate_blocks(&cps[c], 0, nullptr, 0);
		const size_t at = out.size();
		out.resize(at + n);
		gb200_enumerate_blocks(&cps[c], 0, out.data() + at, n);
		for (size_t i = at; i < out.size(); ++i) out[i].compno = c;
	}
}

int decode_init(grk_header_info *h, grk_image *img) {
	using namespace grk;
	DecodeJob &J = *g_job;
	J.status = 1;
	if (!h || !img || !g_ctx) return 1;
	/* what one grk_plugin_tile / this build of the kernels cannot express: the host decodes on its own */
	if (h->cp_tw * h->cp_th != 1 || h->cblk_sty != 0 || h->mct > 1) return 1;
	if (J.reduce >= h->numresolutions) return 1;
	const uint32_t nc = img->numcomps;
	J.nc = nc;
	J.cps.assign(nc, gb200_comp_params());
	for (uint32_t c = 0; c < nc; ++c) {
		gb200_comp_params &cp = J.cps[c];
		memset(&cp, 0, sizeof(cp));
		const grk_image_comp *ic = img->comps + c;
		cp.x0 = ceildiv<uint32_t>(img->x0, ic->dx); cp.y0 = ceildiv<uint32_t>(img->y0, ic->dy);
		cp.x1 = ceildiv<uint32_t>(img->x1, ic->dx); cp.y1 = ceildiv<uint32_t>(img->y1, ic->dy);
		cp.numres = h->numresolutions;
		cp.cblkw_expn = uint_floorlog2(h->cblockw_init); cp.cblkh_expn = uint_floorlog2(h->cblockh_init);
		for (uint32_t r = 0; r < cp.numres; ++r) { cp.prcw_expn[r] = uint_floorlog2(h->prcw_init[r]); cp.prch_expn[r] = uint_floorlog2(h->prch_init[r]); }
		cp.qmfbid = h->irreversible ? 0 : 1;
		cp.prec = ic->prec; cp.sgnd = ic->sgnd;
		cp.dc_shift = ic->sgnd ? 0 : 1 << (ic->prec - 1);
		for (uint32_t b = 0; b < 3 * cp.numres - 2; ++b) { cp.stepsize[b] = 1.0f; cp.inv_step[b] = 8192; cp.band_numbps[b] = 30; cp.rd_weight[b] = 1.0; }
	}
	memset(&J.tp, 0, sizeof(J.tp));
	J.tp.numcomps = nc;
	J.tp.mct = h->mct;
	J.tp.numres_decode = 0; /* the tree spans every resolution; the decode plan is cut to numresolutions - reduce later */
	J.tp.comps = J.cps.data();
	enumerate_image(J.cps, J.table);
	build_tree(J.T, J.table.data(), J.table.size(), nc, J.cps.data());
	/* the host writes each block's bytes without a capacity field (plugin_bridge.cpp:71-78): every block gets the
	 * reference encoder's own worst case (TileProcessor.cpp:2003-2018) plus slack as its slot, and the arena ends with
	 * as many spare bytes as the whole codestream file holds.  No block can carry more bytes than the file, so even a
	 * crafted stream that overfills a slot stays inside the arena; decode_one checks every length against its slot after
	 * the Tier-2 callback and leaves such a stream to the host. */
	const size_t nb = J.T.blocks.size();
	J.offset.resize(nb);
	J.capacity.resize(nb);
	uint64_t total = 0;
	for (size_t i = 0; i < nb; ++i) {
		J.offset[i] = total;
		J.capacity[i] = (J.T.blocks[i].numPix * 4 + 64 + 15) / 16 * 16;
		total += J.capacity[i];
	}
	J.T.data.assign(total + J.file_bytes + 64, 0);
	for (size_t i = 0; i < nb; ++i) J.T.blocks[i].compressedData = J.T.data.data() + J.offset[i];
	J.ready = true;
	J.status = 0;
	return 0;
}

} // namespace

namespace {

/* one codestream through the staged protocol; infile / outfile as the callback shall see them */
/* the host's callback is its own single-threaded code (stream, codec, image store): with several devices working on
 * different codestreams the calls into it are serialised, and the POST_T1 calls (the host stores the image) keep file order */
std::mutex g_host_mu;
struct PostOrder {
	std::mutex mu;
	std::condition_variable cv;
	size_t next = 0;
	void wait_turn(size_t i) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return next == i; }); }
	void done(size_t i) { std::lock_guard<std::mutex> lk(mu); if (next == i) next = i + 1; cv.notify_all(); }
};

int32_t decode_one(gb200_ctx *ctx, grk_decompress_parameters *dp, const std::string &infile, const std::string &outfile, GRK_SUPPORTED_FILE_FMT fmt,
		grk::PLUGIN_DECODE_USER_CALLBACK user_callback, PostOrder *order = nullptr, size_t index = 0) {
	auto callback = [&](grk::PluginDecodeCallbackInfo *i) { std::lock_guard<std::mutex> lk(g_host_mu); return user_callback(i); };
	DecodeJob J;
	J.reduce = dp->core.cp_reduce;
	{
		FILE *f = fopen(infile.c_str(), "rb");
		if (!f) return 2;
		if (fseek(f, 0, SEEK_END) == 0) { const long n = ftell(f); if (n > 0) J.file_bytes = (uint64_t) n; }
		fclose(f);
		if (!J.file_bytes) return 2;
	}
	g_job = &J;
	grk::PluginDecodeCallbackInfo info(infile, outfile, dp, fmt, GRK_DECODE_HEADER);
	info.init_decoders_func = decode_init;
	auto clean = [&](int32_t rc) {
		info.decode_flags = GRK_PLUGIN_DECODE_CLEAN;
		info.init_decoders_func = nullptr;
		try { callback(&info); } catch (...) {}
		g_job = nullptr;
		return rc;
	};
	int32_t rc;
	try { rc = callback(&info); } catch (...) { rc = 7; }
	if (rc || !J.ready) return clean(rc ? rc : 1);
	info.init_decoders_func = nullptr;
	info.tile = &J.T.tile;
	info.decode_flags = GRK_DECODE_T2;
	try { rc = callback(&info); } catch (...) { rc = 7; } /* PluginDecodeUnsupportedException: multi-segment blocks etc. */
	if (rc || !info.image) return clean(rc ? rc : 1);
	/* ---- device decode ---- */
	const size_t nb = J.T.blocks.size();
	for (size_t k = 0; k < nb; ++k) /* a block that outgrew its slot has overwritten its neighbours: not decodable from this arena */
		if (J.T.blocks[k].compressedDataLength > J.capacity[k]) return clean(8);
	for (uint32_t c = 0; c < J.nc; ++c) {
		const grk_plugin_tile_component *tc = J.T.tile.tileComponents[c];
		for (uint32_t r = 0; r < tc->numResolutions; ++r)
			for (uint32_t b = 0; b < tc->resolutions[r]->numBands; ++b)
				J.cps[c].stepsize[r ? 3 * r - 2 + b : 0] = tc->resolutions[r]->bands[b]->stepsize; /* carries the decoder's x0.5 */
	}
	J.tp.numres_decode = J.cps[0].numres - J.reduce;
	gb200_plan *plan = cache_of(ctx)->get(ctx, J.tp, false); /* owned by the context's cache */
	if (!plan) { say("gb200_plan_create"); return clean(3); }
	const size_t nbd = gb200_plan_num_blocks(plan); /* blocks of the resolutions that are reconstructed */
	const gb200_cblk_info *dinfo = gb200_plan_blocks(plan);
	std::vector<gb200_cblk_dec> in(nbd);
	{ /* the reduced table is the full one minus the blocks of the dropped resolutions, same order */
		size_t i = 0;
		const gb200_cblk_info *finfo = J.table.data();
		for (size_t k = 0; k < nb && i < nbd; ++k) {
			if (finfo[k].compno != dinfo[i].compno || finfo[k].resno != dinfo[i].resno || finfo[k].bandno != dinfo[i].bandno
					|| finfo[k].precno != dinfo[i].precno || finfo[k].cblkno != dinfo[i].cblkno) continue;
			const grk_plugin_code_block &B = J.T.blocks[k];
			memset(&in[i], 0, sizeof(in[i]));
			in[i].numbps = (uint32_t) B.numBitPlanes;
			in[i].numpasses = (uint32_t) B.numPasses;
			in[i].data_len = (uint32_t) B.compressedDataLength;
			in[i].data_offset = J.offset[k];
			++i;
		}
		if (i != nbd) return clean(5);
	}
	std::vector<int32_t*> planes(J.nc);
	for (uint32_t c = 0; c < J.nc; ++c) {
		grk_image_comp *ic = info.image->comps + c;
		const uint32_t top = J.reduce;
		const uint32_t w = grk::uint_ceildivpow2(J.cps[c].x1, top) - grk::uint_ceildivpow2(J.cps[c].x0, top);
		const uint32_t hgt = grk::uint_ceildivpow2(J.cps[c].y1, top) - grk::uint_ceildivpow2(J.cps[c].y0, top);
		if (ic->w != w || ic->h != hgt) return clean(5);
		if (!ic->data && !grk_image_single_component_data_alloc(ic)) return clean(5);
		planes[c] = ic->data;
	}
	if (gb200_decode_tiles(plan, in.data(), J.T.data.data(), J.T.data.size() - 64 - J.file_bytes, planes.data()) != GB200_OK) { say("gb200_decode_tiles"); return clean(4); }
	g_decodes++;
	for (size_t k = 0; k < g_ctxs.size() && k < 16; ++k) if (g_ctxs[k] == ctx) g_per_device[k]++;
	if (order) order->wait_turn(index); /* the host stores the images in file order */
	info.decode_flags = GRK_DECODE_POST_T1;
	try { rc = callback(&info); } catch (...) { rc = 7; }
	return clean(rc);
}

} // namespace

extern "C" PLUGIN_API int32_t plugin_decode(grk_decompress_parameters *dp, grk::PLUGIN_DECODE_USER_CALLBACK callback) {
	if (!g_ctx || !dp || !callback) return -1;
	return decode_one(g_ctx, dp, dp->infile, dp->outfile, dp->decod_format, callback);
}

/* ---- batch decode: plugin_init_batch_decode / plugin_batch_decode / plugin_stop_batch_decode (plugin_interface.h:124-130).
 * The shipped CLI calls plugin_batch_decode only when the init call FAILED (`if (success) success = ...batch_decode()`,
 * grk_decompress.cpp:1242-1247) and then polls plugin_is_batch_complete, so the work starts in the init call; a host that
 * does call plugin_batch_decode afterwards finds the batch already running.  One thread walks the .j2k / .j2c / .jp2 files
 * of input_dir in name order and runs each through the staged protocol; a stream the ABI cannot express is left to the
 * host by passing GRK_DECODE_ALL (full CPU decode inside the callback, grk_decompress.cpp:1338-1342). */
namespace {

struct DecBatch {
	std::vector<std::thread> workers; /* one per context (device): codestream i is decoded on device i mod N */
	PostOrder order;
	std::atomic<size_t> finished{0};
	std::vector<std::string> files;
	std::string out_dir;
	grk_decompress_parameters params;
	grk::PLUGIN_DECODE_USER_CALLBACK callback = nullptr;
	std::atomic<bool> stop{false}, complete{true};
};
DecBatch *g_dec_batch = nullptr;

const char *out_extension(uint32_t cod_format) {
	switch (cod_format) {
	case GRK_PXM_FMT: return "ppm";
	case GRK_PGX_FMT: return "pgx";
	case GRK_BMP_FMT: return "bmp";
	case GRK_TIF_FMT: return "tif";
	case GRK_RAW_FMT: return "raw";
	case GRK_RAWL_FMT: return "rawl";
	case GRK_PNG_FMT: return "png";
	case GRK_JPG_FMT: return "jpg";
	default: return "out";
	}
}

} // namespace

extern "C" PLUGIN_API int32_t plugin_init_batch_decode(const char *input_dir, const char *output_dir, grk_decompress_parameters *dp,
		grk::PLUGIN_DECODE_USER_CALLBACK callback) {
	if (!g_ctx || !input_dir || !output_dir || !dp || !callback) return -1;
	if (g_dec_batch) {
		if (!g_dec_batch->complete) return -1;
		for (auto &t : g_dec_batch->workers) if (t.joinable()) t.join();
		delete g_dec_batch;
		g_dec_batch = nullptr;
	}
	DIR *d = opendir(input_dir);
	if (!d) return 2;
	DecBatch *b = new DecBatch();
	while (dirent *e = readdir(d)) {
		std::string n = e->d_name;
		const size_t dot = n.rfind('.');
		if (dot == std::string::npos) continue;
		std::string ext = n.substr(dot + 1);
		for (auto &c : ext) c = (char) tolower(c);
		if (ext == "j2k" || ext == "j2c" || ext == "jp2") b->files.push_back(n);
	}
	closedir(d);
	std::sort(b->files.begin(), b->files.end());
	b->out_dir = output_dir;
	b->params = *dp;
	b->callback = callback;
	b->complete = false;
	g_dec_batch = b;
	const std::string in_dir = input_dir;
	const size_t ndev = g_ctxs.size();
	if (b->files.empty()) b->complete = true;
	for (size_t k = 0; k < ndev; ++k)
		b->workers.emplace_back([b, in_dir, k, ndev]() {
			for (size_t i = k; i < b->files.size(); i += ndev) {
				const std::string &n = b->files[i];
				if (!b->stop) {
					const size_t dot = n.rfind('.');
					std::string ext = n.substr(dot + 1);
					for (auto &c : ext) c = (char) tolower(c);
					const GRK_SUPPORTED_FILE_FMT fmt = ext == "jp2" ? GRK_JP2_FMT : GRK_J2K_FMT;
					const std::string in = in_dir + "/" + n, out = b->out_dir + "/" + n.substr(0, dot) + "." + out_extension(b->params.cod_format);
					grk_decompress_parameters fp = b->params;
					fp.infile[0] = 0; fp.outfile[0] = 0; /* the callback takes the names from the info struct */
					fp.decod_format = fmt;
					/* decode_one waits for this codestream's turn in the file order before the host stores the image; a stream
					 * the ABI cannot express is decoded by the host on its own, in one callback, when its turn has come */
					const int32_t rc = decode_one(g_ctxs[k], &fp, in, out, fmt, b->callback, &b->order, i);
					b->order.wait_turn(i);
					if (rc) {
						grk::PluginDecodeCallbackInfo info(in, out, &fp, fmt, GRK_DECODE_ALL);
						std::lock_guard<std::mutex> lk(g_host_mu);
						try { b->callback(&info); } catch (...) {}
					}
					b->order.done(i);
				} else { b->order.wait_turn(i); b->order.done(i); }
				if (++b->finished == b->files.size()) b->complete = true;
			}
		});
	return 0;
}

bool decode_batch_complete() { return !g_dec_batch || g_dec_batch->complete; }
extern "C" PLUGIN_API int32_t plugin_batch_decode(void) { return g_dec_batch ? 0 : -1; }

extern "C" PLUGIN_API void plugin_stop_batch_decode(void) {
	if (!g_dec_batch) return;
	g_dec_batch->stop = true;
	for (auto &t : g_dec_batch->workers) if (t.joinable()) t.join();
	delete g_dec_batch;
	g_dec_batch = nullptr;
}

extern "C" PLUGIN_API uint32_t plugin_get_debug_state(void) { return GRK_PLUGIN_STATE_NO_DEBUG; }
extern "C" PLUGIN_API void plugin_debug_mqc_next_cxd(grk::grk_plugin_debug_mqc *, uint32_t) {}
extern "C" PLUGIN_API void plugin_debug_mqc_next_plane(grk::grk_plugin_debug_mqc *) {}
