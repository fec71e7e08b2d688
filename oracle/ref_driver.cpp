/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain C ABI around the UNMODIFIED reference library (oracle/_ref/libgrok_ref.so,
 * built by oracle/Makefile.ref from /root/reference).  It lets the Python tests call the
 * reference's own stage functions on raw buffers so that the C restatement (gb_oracle.c)
 * and the CUDA kernels can be pinned against the real thing:
 *
 *   ref_mct_*          -> grk::mct::{encode,decode}_{rev,irrev}     mct/mct.cpp:85,143,195,352
 *   ref_dwt_encode     -> grk::Wavelet::encode                      transform/Wavelet.cpp:35
 *   ref_dwt_decode     -> grk::Wavelet::decode                      transform/Wavelet.cpp, dwt.cpp:1208,2154
 *   ref_t1_encode_cblk -> grk::t1_encode_cblk                       t1/t1_part1/t1.cpp:1182
 *   ref_t1_decode_cblk -> grk::t1_decode_cblk                       t1/t1_part1/t1.cpp:1038
 *   ref_dwt_norm       -> grk::dwt_utils::getnorm_{53,97}           transform/dwt_utils.cpp:143
 *   ref_band_stepsize  -> grk::Quantizer::setBandStepSizeAndBps     codestream/Quantizer.cpp:65
 *   ref_qcd_generate   -> grk::param_qcd::generate                  codestream/HTParams.cpp:164
 *
 * Only the reference's headers are included (a build-time include dependency); no reference
 * source is copied here.
 */
#include "ojph_block_encoder.h" /* before grok_includes.h, which poisons malloc / free (as T1HT.cpp does) */
#include "ojph_block_decoder.h"
#include "ojph_mem.h"
#include "grok_includes.h"
#include "t1_common.h"
#include "Tier1.h"
#include "dwt53.h"
#include "dwt97.h"
#include "HTParams.h"
#include "RateControl.h"
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>
#include <unistd.h>
#include <execinfo.h>
#include <signal.h>

using namespace grk;

static inline uint32_t cdp2(uint32_t a, uint32_t b) { return (uint32_t) (((uint64_t) a + ((uint64_t) 1 << b) - 1) >> b); }

extern "C" {

int ref_init(uint32_t nthreads) {
	return grk_initialize(nullptr, nthreads) ? 0 : 1;
}

void ref_mct_encode_rev(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { mct::encode_rev(c0, c1, c2, n); }
void ref_mct_decode_rev(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { mct::decode_rev(c0, c1, c2, n); }
void ref_mct_encode_irrev(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n) { mct::encode_irrev(c0, c1, c2, n); }
void ref_mct_decode_irrev(float *c0, float *c1, float *c2, uint64_t n) { mct::decode_irrev(c0, c1, c2, n); }

double ref_dwt_norm(uint32_t level, uint32_t orient, int reversible) {
	return reversible ? dwt_utils::getnorm_53(level, (uint8_t) orient) : dwt_utils::getnorm_97(level, (uint8_t) orient);
}
double ref_mct_norm(uint32_t compno, int reversible) {
	return reversible ? mct::get_norms_rev()[compno] : mct::get_norms_irrev()[compno];
}

/* A TileComponent carrying just what Wavelet::{encode,decode} read: the resolution rectangles,
 * the tile-component rectangle and a TileBuffer that aliases the caller's plane. */
struct FakeTilec {
	TileComponent *tc;
	FakeTilec(int32_t *data, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
			uint32_t numres_decode, bool encoder) {
		tc = new TileComponent();
		tc->numresolutions = numres;
		tc->numAllocatedResolutions = numres;
		tc->minimum_num_resolutions = numres_decode;
		tc->m_is_encoder = encoder;
		tc->resolutions = new grk_tcd_resolution[numres];
		for (uint32_t r = 0; r < numres; ++r) {
			uint32_t lvl = numres - 1 - r;
			auto res = tc->resolutions + r;
			res->x0 = cdp2(x0, lvl);
			res->y0 = cdp2(y0, lvl);
			res->x1 = cdp2(x1, lvl);
			res->y1 = cdp2(y1, lvl);
			for (int b = 0; b < 3; ++b)
				res->bands[b].precincts = nullptr;
		}
		auto top = tc->resolutions + (encoder ? numres : numres_decode) - 1;
		tc->x0 = top->x0; tc->y0 = top->y0; tc->x1 = top->x1; tc->y1 = top->y1;
		tc->buf = new TileBuffer();
		tc->buf->data = data;
		tc->buf->owns_data = false;
		tc->buf->data_size = 0;
		tc->buf->data_size_needed = 0;
		tc->buf->reduced_image_dim = grk_rect(top->x0, top->y0, top->x1, top->y1);
		tc->buf->reduced_tile_dim = tc->buf->reduced_image_dim;
	}
	~FakeTilec() {
		tc->buf->data = nullptr;
		delete tc;
	}
};

/* in place on `data` (row stride = x1-x0), tile-component canvas rectangle [x0,x1)x[y0,y1) */
int ref_dwt_encode(int32_t *data, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
		int qmfbid) {
	FakeTilec f(data, x0, y0, x1, y1, numres, numres, true);
	Wavelet w;
	return w.encode(f.tc, (uint8_t) qmfbid) ? 0 : 1;
}

/* in place; row stride = width of resolution numres_decode-1 (how -r works, dwt.cpp:735) */
int ref_dwt_decode(int32_t *data, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t numres,
		uint32_t numres_decode, int qmfbid) {
	FakeTilec f(data, x0, y0, x1, y1, numres, numres_decode, false);
	/* Wavelet::decode only reads p_tcd->whole_tile_decoding */
	void *mem = grk_calloc(1, sizeof(TileProcessor));
	auto tp = reinterpret_cast<TileProcessor*>(mem);
	tp->whole_tile_decoding = true;
	Wavelet w;
	bool ok = w.decode(tp, f.tc, numres_decode, (uint8_t) qmfbid);
	grok_free(mem);
	return ok ? 0 : 1;
}

/* Tier-1 encode of one code block.  `data` holds w*h quantised coefficients WITH the 6 fractional
 * bits (what T1Part1::preEncode produces).  Returns total passes; fills numbps, per-pass rate/len/
 * distortion, and the byte stream (out must hold >= w*h*4+16 bytes). */
static uint8_t *g_terms_out = nullptr; /* optional: per-pass termination flags of the next ref_t1_encode_cblk call */
void ref_t1_want_terms(uint8_t *terms) { g_terms_out = terms; }

int ref_t1_encode_cblk(const int32_t *data, uint32_t w, uint32_t h, uint32_t orient, uint32_t compno,
		uint32_t level, uint32_t qmfbid, double stepsize, uint32_t cblksty, const double *mct_norms,
		uint32_t mct_numcomps, int do_rate_control, uint8_t *out, uint32_t *numbps, uint32_t *rates,
		uint32_t *lens, double *dists, double *total_dist) {
	t1_info *t1 = t1_create(true);
	if (!t1 || !t1_allocate_buffers(t1, w, h))
		return -1;
	t1->data_stride = w;
	uint32_t mx = 0;
	for (uint32_t i = 0; i < w * h; ++i) {
		t1->data[i] = data[i];
		uint32_t a = (uint32_t) abs(data[i]);
		if (a > mx) mx = a;
	}
	/* reference allocates 2 zeroed pad bytes in front (TileProcessor.cpp:1997-2018) */
	size_t cap = (size_t) w * h * 4 + 1024; /* slack for the flush bytes of terminated passes on tiny blocks */
	uint8_t *buf = (uint8_t*) grk_calloc(1, cap);
	tcd_cblk_enc_t cblk;
	memset(&cblk, 0, sizeof(cblk));
	cblk.x0 = 0; cblk.y0 = 0; cblk.x1 = w; cblk.y1 = h;
	cblk.data = buf + 2;
	cblk.data_size = (uint32_t) (cap - 2);
	double d = t1_encode_cblk(t1, &cblk, mx, (uint8_t) orient, compno, level, qmfbid, stepsize, cblksty,
			mct_norms, mct_numcomps, do_rate_control != 0);
	*numbps = cblk.numbps;
	int np = (int) cblk.totalpasses;
	for (int i = 0; i < np; ++i) {
		rates[i] = cblk.passes[i].rate;
		lens[i] = cblk.passes[i].len;
		dists[i] = cblk.passes[i].distortiondec;
		if (g_terms_out) g_terms_out[i] = (uint8_t) cblk.passes[i].term;
	}
	g_terms_out = nullptr;
	if (np > 0)
		memcpy(out, cblk.data, rates[np - 1]);
	if (total_dist) *total_dist = d;
	t1_code_block_enc_deallocate(&cblk);
	grok_free(buf);
	t1_destroy(t1);
	return np;
}

/* Tier-1 decode of one single-segment code block; out gets t1->data (values still carry the
 * extra low bit; T1Part1::post_decode halves them). */
int ref_t1_decode_cblk(const uint8_t *bytes, uint32_t len, uint32_t numpasses, uint32_t numbps,
		uint32_t orient, uint32_t roishift, uint32_t cblksty, uint32_t w, uint32_t h, int32_t *out) {
	t1_info *t1 = t1_create(false);
	if (!t1)
		return -1;
	uint8_t *buf = (uint8_t*) grk_calloc(1, (size_t) len + 16);
	memcpy(buf, bytes, len);
	tcd_seg_data_chunk_t chunk;
	chunk.data = buf;
	chunk.len = len + GRK_FAKE_MARKER_BYTES;
	tcd_seg_t seg;
	memset(&seg, 0, sizeof(seg));
	seg.len = len;
	seg.real_num_passes = numpasses;
	tcd_cblk_dec_t cblk;
	memset(&cblk, 0, sizeof(cblk));
	cblk.numchunks = 1;
	cblk.chunks = &chunk;
	cblk.x0 = 0; cblk.y0 = 0; cblk.x1 = w; cblk.y1 = h;
	cblk.real_num_segs = 1;
	cblk.segs = &seg;
	cblk.numbps = numbps;
	bool ok = t1_decode_cblk(t1, &cblk, orient, roishift, cblksty, false);
	if (ok)
		memcpy(out, t1->data, (size_t) w * h * sizeof(int32_t));
	grok_free(buf);
	t1_destroy(t1);
	return ok ? 0 : 1;
}

/* Tier-1 decode of a code block given as codeword segments (what Tier-2 delivers with TERMALL / LAZY) */
int ref_t1_decode_cblk_segs(const uint8_t *bytes, const uint32_t *seg_len, const uint32_t *seg_passes, uint32_t nsegs,
		uint32_t numbps, uint32_t orient, uint32_t roishift, uint32_t cblksty, uint32_t w, uint32_t h, int32_t *out) {
	t1_info *t1 = t1_create(false);
	if (!t1)
		return -1;
	uint64_t total = 0;
	for (uint32_t i = 0; i < nsegs; ++i) total += seg_len[i];
	uint8_t *buf = (uint8_t*) grk_calloc(1, (size_t) total + 16);
	memcpy(buf, bytes, total);
	tcd_seg_data_chunk_t chunk;
	chunk.data = buf;
	chunk.len = (uint32_t) total + GRK_FAKE_MARKER_BYTES;
	std::vector<tcd_seg_t> segs(nsegs ? nsegs : 1);
	memset(segs.data(), 0, segs.size() * sizeof(tcd_seg_t));
	for (uint32_t i = 0; i < nsegs; ++i) { segs[i].len = seg_len[i]; segs[i].real_num_passes = seg_passes[i]; }
	tcd_cblk_dec_t cblk;
	memset(&cblk, 0, sizeof(cblk));
	cblk.numchunks = 1;
	cblk.chunks = &chunk;
	cblk.x0 = 0; cblk.y0 = 0; cblk.x1 = w; cblk.y1 = h;
	cblk.real_num_segs = nsegs;
	cblk.segs = segs.data();
	cblk.numbps = numbps;
	bool ok = t1_decode_cblk(t1, &cblk, orient, roishift, cblksty, false);
	if (ok)
		memcpy(out, t1->data, (size_t) w * h * sizeof(int32_t));
	grok_free(buf);
	t1_destroy(t1);
	return ok ? 0 : 1;
}

/* default quantisation: exponent/mantissa per band as grk_compress derives them
 * (j2k.cpp:1839-1841 -> HTParams.cpp:164).  expn/mant have 3*decomps+1 entries. */
void ref_qcd_generate(uint32_t guard_bits, uint32_t decomps, int reversible, uint32_t prec,
		int color_transform, int is_signed, uint32_t *expn, uint32_t *mant) {
	param_qcd q;
	q.generate((uint8_t) guard_bits, decomps, reversible != 0, prec, color_transform != 0, is_signed != 0);
	grk_stepsize steps[GRK_J2K_MAXBANDS];
	memset(steps, 0, sizeof(steps));
	q.pull(steps, reversible != 0);
	for (uint32_t i = 0; i < 3 * decomps + 1; ++i) {
		expn[i] = steps[i].expn;
		mant[i] = steps[i].mant;
	}
}

/* band stepsize / numbps / inv_step exactly as Quantizer::setBandStepSizeAndBps computes them */
void ref_band_stepsize(uint32_t expn, uint32_t mant, uint32_t resno, uint32_t bandno_in_res, uint32_t orient,
		int qmfbid, uint32_t numgbits, uint32_t prec, float fraction, float *stepsize, uint32_t *numbps,
		uint32_t *inv_step) {
	grk_tcp tcp;
	tcp.isHT = false;
	grk_tccp tccp;
	memset(&tccp, 0, sizeof(tccp));
	tccp.qmfbid = (uint8_t) qmfbid;
	tccp.numgbits = (uint8_t) numgbits;
	tccp.roishift = 0;
	uint32_t offset = (resno == 0) ? 0 : 3 * resno - 2;
	tccp.stepsizes[offset + bandno_in_res].expn = (uint8_t) expn;
	tccp.stepsizes[offset + bandno_in_res].mant = (uint16_t) mant;
	grk_tcd_band band;
	band.bandno = (uint8_t) orient;
	Quantizer q;
	q.setBandStepSizeAndBps(&tcp, &band, resno, (uint8_t) bandno_in_res, &tccp, prec, fraction);
	*stepsize = band.stepsize;
	*numbps = band.numbps;
	*inv_step = band.inv_step;
}

/* ---- whole codec through the reference's public API (grok.h), memory streams only ------------ */

static void quiet_cb(const char *, void *) {}

/* ---- HTJ2K block coder (t1/t1_ht/coding): the cleanup-pass encoder and decoder on one block of sign-magnitude samples, the form
 * T1HT::preEncode hands them over in (T1HT.cpp:56-133, 135-175) */
int ref_ht_encode_block(const int32_t *sm, int w, int h, int stride, int missing_msbs, uint8_t *out, int cap) {
	ojph::mem_elastic_allocator elastic(1048576);
	ojph::coded_lists *coded = nullptr;
	int lengths[2] = {0, 0};
	ojph::local::ojph_encode_codeblock((ojph::si32*) sm, missing_msbs, 1, w, h, stride, lengths, &elastic, coded);
	if (!coded || lengths[0] > cap) return -1;
	memcpy(out, coded->buf, (size_t) lengths[0]);
	return lengths[0];
}
int ref_ht_decode_block(const uint8_t *data, int len, int missing_msbs, int w, int h, int stride, int32_t *out) {
	/* the decoder moves 32 bits at a time and may touch a few bytes either side of the segment */
	std::vector<uint8_t> padded((size_t) len + 64, 0);
	memcpy(padded.data() + 32, data, (size_t) len);
	ojph::local::ojph_decode_codeblock(padded.data() + 32, (ojph::si32*) out, missing_msbs, 1, len, 0, w, h, stride);
	return 0;
}

/* code-block style byte (grk_compress -M) applied by the following ref_encode_image / ref_plugin_encode_file calls */
static uint32_t g_cblk_sty = 0;
/* RateControl::convexHull (t2/RateControl.cpp:31) on one block's pass table: len / cumulative distortion in, log slopes out */
void ref_rd_convex_hull(const uint32_t *len, const double *dist, uint32_t numpasses, uint16_t *slope) {
	std::vector<grk_tcd_pass> passes(numpasses ? numpasses : 1);
	for (uint32_t p = 0; p < numpasses; ++p) { passes[p].len = len[p]; passes[p].distortiondec = dist[p]; }
	RateControl::convexHull(passes.data(), numpasses);
	for (uint32_t p = 0; p < numpasses; ++p) slope[p] = passes[p].slope;
}

void ref_set_cblk_sty(uint32_t sty) { g_cblk_sty = sty; }
/* max-shift region of interest (grk_compress -ROI c=compno,U=shift) for the following ref_encode_image calls; compno < 0 = none */
static int32_t g_roi_compno = -1;
static uint32_t g_roi_shift = 0;
void ref_set_roi(int32_t compno, uint32_t shift) { g_roi_compno = compno; g_roi_shift = shift; }

/* precinct sizes (grk_compress -c [w,h],[w,h],...: first entry = highest resolution, the last one is halved for every further
 * resolution, j2k.cpp:2001-2048) for the following ref_encode_image calls; n = 0: the default (maximal precincts) */
static uint32_t g_prc_n = 0, g_prc_w[GRK_J2K_MAXRLVLS], g_prc_h[GRK_J2K_MAXRLVLS];
void ref_set_precincts(uint32_t n, const uint32_t *w, const uint32_t *h) {
	g_prc_n = n > GRK_J2K_MAXRLVLS ? GRK_J2K_MAXRLVLS : n;
	for (uint32_t i = 0; i < g_prc_n; ++i) { g_prc_w[i] = w[i]; g_prc_h[i] = h[i]; }
}
/* progression order (grk_compress -p: 0 LRCP, 1 RLCP, 2 RPCL, 3 PCRL, 4 CPRL) for the following ref_encode_image calls; < 0: default */
static int g_prog = -1;
void ref_set_progression(int prog) { g_prog = prog; }

/* grk_compress-equivalent: planar int32 image -> raw J2K codestream.
 *   tile_w/tile_h 0 = single tile;  rates[numlayers] = compression ratios (-r); numlayers 0 = lossless
 *   cinema2k_fps 24/48 = -w profile.  Returns the codestream length, or -1. */
int64_t ref_encode_image(uint32_t numcomps, uint32_t w, uint32_t h, uint32_t prec, uint32_t sgnd,
		const int32_t *const *planes, uint32_t tile_w, uint32_t tile_h, uint32_t numres, uint32_t cblkw, uint32_t cblkh,
		int irreversible, uint32_t numlayers, const double *rates, int cinema2k_fps, uint32_t rc_algorithm,
		uint8_t *out, uint64_t cap) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	grk_cparameters param;
	grk_set_default_encoder_parameters(&param);
	param.numresolution = numres;
	param.cblockw_init = cblkw;
	param.cblockh_init = cblkh;
	param.irreversible = irreversible != 0;
	param.rateControlAlgorithm = rc_algorithm;
	param.cblk_sty = (uint8_t) g_cblk_sty;
	param.isHT = (g_cblk_sty & GRK_CBLKSTY_HT) != 0; /* grk_compress -M 64 (grk_compress.cpp:1131-1141) */
	param.roi_compno = g_roi_compno;
	param.roi_shift = g_roi_compno >= 0 ? g_roi_shift : 0;
	if (g_prc_n) {
		param.csty |= 0x01;
		param.res_spec = g_prc_n;
		for (uint32_t i = 0; i < g_prc_n; ++i) { param.prcw_init[i] = g_prc_w[i]; param.prch_init[i] = g_prc_h[i]; }
	}
	if (g_prog >= 0) param.prog_order = (GRK_PROG_ORDER) g_prog;
	if (tile_w && tile_h) {
		param.tile_size_on = true;
		param.cp_tdx = tile_w;
		param.cp_tdy = tile_h;
	}
	if (numlayers) {
		param.tcp_numlayers = numlayers;
		for (uint32_t i = 0; i < numlayers; ++i) param.tcp_rates[i] = rates[i];
		param.cp_disto_alloc = 1;
	} else {
		/* grk_compress.cpp: no -r/-q given => one lossless layer */
		param.tcp_numlayers = 1;
		param.tcp_rates[0] = 0;
		param.cp_disto_alloc = 1;
	}
	if (cinema2k_fps) {
		param.rsiz = GRK_PROFILE_CINEMA_2K;
		param.framerate = cinema2k_fps;
		if (cinema2k_fps == 24) { param.max_cs_size = GRK_CINEMA_24_CS; param.max_comp_size = GRK_CINEMA_24_COMP; }
		else { param.max_cs_size = GRK_CINEMA_48_CS; param.max_comp_size = GRK_CINEMA_48_COMP; }
	}
	param.tcp_mct = (numcomps >= 3) ? 1 : 0; /* grk_compress.cpp:1997-1998 */
	std::vector<grk_image_cmptparm> cp(numcomps);
	for (uint32_t i = 0; i < numcomps; ++i) {
		memset(&cp[i], 0, sizeof(grk_image_cmptparm));
		cp[i].dx = cp[i].dy = 1;
		cp[i].w = w; cp[i].h = h;
		cp[i].prec = prec; cp[i].sgnd = sgnd;
	}
	grk_image *image = grk_image_create(numcomps, cp.data(), numcomps >= 3 ? GRK_CLRSPC_SRGB : GRK_CLRSPC_GRAY);
	if (!image) return -1;
	image->x0 = 0; image->y0 = 0; image->x1 = w; image->y1 = h;
	for (uint32_t i = 0; i < numcomps; ++i)
		memcpy(image->comps[i].data, planes[i], (size_t) w * h * sizeof(int32_t));
	int64_t len = -1;
	grk_stream *stream = grk_stream_create_mem_stream(out, cap, false, false);
	grk_codec *codec = stream ? grk_create_compress(GRK_CODEC_J2K, stream) : nullptr;
	if (codec && grk_setup_encoder(codec, &param, image) && grk_start_compress(codec, image)
			&& grk_encode(codec) && grk_end_compress(codec))
		len = (int64_t) grk_stream_get_write_mem_stream_length(stream);
	if (stream) grk_stream_destroy(stream);
	if (codec) grk_destroy_codec(codec);
	grk_image_destroy(image);
	return len;
}

/* decode window (grk_decompress -d x0,y0,x1,y1) for the following ref_decode_image calls; all zero = whole image */
static uint32_t g_da[4] = {0, 0, 0, 0};
void ref_set_decode_area(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) { g_da[0] = x0; g_da[1] = y0; g_da[2] = x1; g_da[3] = y1; }

/* grk_decompress-equivalent. planes_out[c] must hold the (reduced) component; returns 0 on success and
 * fills dims[0..1] = decoded width/height, dims[2] = numcomps. */
int ref_decode_image(const uint8_t *buf, uint64_t len, uint32_t reduce, uint32_t layers, int32_t *const *planes_out,
		uint64_t plane_capacity, uint32_t *dims) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	grk_dparameters dp;
	grk_set_default_decoder_parameters(&dp);
	dp.cp_reduce = reduce;
	dp.cp_layer = layers;
	grk_stream *stream = grk_stream_create_mem_stream(const_cast<uint8_t*>(buf), len, false, true);
	grk_codec *codec = stream ? grk_create_decompress(GRK_CODEC_J2K, stream) : nullptr;
	grk_image *image = nullptr;
	int rc = 1;
	if (codec && grk_setup_decoder(codec, &dp) && grk_read_header(codec, nullptr, &image)
			&& ((g_da[2] == 0 && g_da[3] == 0) || grk_set_decode_area(codec, image, g_da[0], g_da[1], g_da[2], g_da[3]))
			&& grk_decode(codec, nullptr, image) && grk_end_decompress(codec)) {
		rc = 0;
		dims[0] = image->comps[0].w; dims[1] = image->comps[0].h; dims[2] = image->numcomps;
		for (uint32_t c = 0; c < image->numcomps; ++c) {
			uint64_t n = (uint64_t) image->comps[c].w * image->comps[c].h;
			if (n > plane_capacity || !image->comps[c].data) { rc = 2; break; }
			memcpy(planes_out[c], image->comps[c].data, n * sizeof(int32_t));
		}
	}
	if (stream) grk_stream_destroy(stream);
	if (codec) grk_destroy_codec(codec);
	if (image) grk_image_destroy(image);
	return rc;
}

/* ---- the official plugin path: grk_plugin_load / init / encode, exactly what `grk_compress -g <dir>` does
 * (grk_compress.cpp:2205-2301), with a callback that mirrors plugin_compress_callback (grk_compress.cpp:1770-2160)
 * but writes the codestream into memory.  Returns the codestream length; -1 = plugin not loaded / init failed
 * (the CLI would fall back to the CPU), -2 = plugin_encode returned non-zero, -3 = the host side failed. */
static uint8_t *g_cb_out = nullptr;
static uint64_t g_cb_cap = 0;
static int64_t g_cb_len = -1;

static bool plugin_cb(grk_plugin_encode_user_callback_info *info) {
	grk_cparameters *param = info->encoder_parameters;
	grk_image *image = info->image;
	g_cb_len = -3;
	if (!image || !info->tile) return false;
	if (param->tcp_mct == 255) param->tcp_mct = (image->numcomps >= 3) ? 1 : 0; /* grk_compress.cpp:1996-1998 */
	grk_stream *stream = grk_stream_create_mem_stream(g_cb_out, g_cb_cap, false, false);
	grk_codec *codec = stream ? grk_create_compress(GRK_CODEC_J2K, stream) : nullptr;
	if (codec && grk_setup_encoder(codec, param, image) && grk_start_compress(codec, image)
			&& grk_encode_with_plugin(codec, info->tile) && grk_end_compress(codec))
		g_cb_len = (int64_t) grk_stream_get_write_mem_stream_length(stream);
	if (stream) grk_stream_destroy(stream);
	if (codec) grk_destroy_codec(codec);
	return g_cb_len >= 0;
}

int64_t ref_plugin_encode_file(const char *plugin_dir, const char *infile, uint32_t tile_w, uint32_t tile_h, uint32_t numres,
		uint32_t cblkw, uint32_t cblkh, int irreversible, uint32_t numlayers, const double *rates, uint32_t rc_algorithm,
		uint8_t *out, uint64_t cap) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	grk_plugin_load_info li;
	li.plugin_path = plugin_dir;
	if (!grk_plugin_load(li)) return -1;
	grk_plugin_init_info ii;
	ii.deviceId = 0;
	ii.verbose = true;
	if (!grk_plugin_init(ii)) { grk_plugin_cleanup(); return -1; }
	grk_cparameters param;
	grk_set_default_encoder_parameters(&param);
	strncpy(param.infile, infile, sizeof(param.infile) - 1);
	param.decod_format = GRK_PXM_FMT;
	param.cod_format = GRK_J2K_FMT;
	param.numresolution = numres;
	param.cblockw_init = cblkw;
	param.cblockh_init = cblkh;
	param.irreversible = irreversible != 0;
	param.rateControlAlgorithm = rc_algorithm;
	if (tile_w && tile_h) { param.tile_size_on = true; param.cp_tdx = tile_w; param.cp_tdy = tile_h; }
	param.tcp_numlayers = numlayers ? numlayers : 1;
	for (uint32_t i = 0; i < numlayers; ++i) param.tcp_rates[i] = rates[i];
	if (!numlayers) param.tcp_rates[0] = 0;
	param.cp_disto_alloc = 1;
	param.tcp_mct = 255; /* "not set on the command line": decided from the component count, grk_compress.cpp:1996-1998 */
	g_cb_out = out; g_cb_cap = cap; g_cb_len = -1;
	int32_t rc = grk_plugin_encode(&param, plugin_cb);
	int64_t len = rc ? -2 : g_cb_len;
	grk_plugin_cleanup();
	return len;
}

/* ---- the official plugin path, decode: grk_plugin_load / init / decode with a callback that mirrors decode_callback /
 * pre_decode / post_decode of grk_decompress.cpp:1336-1560, reading the codestream from memory and handing the pixels
 * back in memory instead of writing a file.  Returns 0; -1 = plugin not loaded / init failed (CLI: CPU path), -2 =
 * plugin_decode returned non-zero (CLI: CPU path), -3 = host side failed. */
static const uint8_t *g_dec_in = nullptr;
static uint64_t g_dec_in_len = 0;
static int32_t *const *g_dec_out = nullptr;
static uint64_t g_dec_cap = 0;
static uint32_t *g_dec_dims = nullptr;
static int g_dec_stored = 0;
static bool g_dec_from_files = false; /* batch decode: input named by the plugin, planes of frame k appended at k * 3 * cap */
static int32_t *g_dec_batch_out = nullptr;
static uint32_t g_dec_batch_max = 0;

static int32_t plugin_dec_cb(grk_plugin_decode_callback_info *info) {
	int32_t rc = -1;
	grk_decompress_parameters *param = info->decoder_parameters;
	if (info->decode_flags & GRK_DECODE_T1) info->init_decoders_func = nullptr;
	if (info->decode_flags & GRK_PLUGIN_DECODE_CLEAN) {
		if (info->l_stream) grk_stream_destroy(info->l_stream);
		info->l_stream = nullptr;
		if (info->l_codec) grk_destroy_codec(info->l_codec);
		info->l_codec = nullptr;
		if (info->image && !info->plugin_owns_image) { grk_image_destroy(info->image); info->image = nullptr; }
		rc = 0;
	}
	if (info->decode_flags & (GRK_DECODE_HEADER | GRK_DECODE_T1 | GRK_DECODE_T2)) { /* pre_decode */
		bool failed = false;
		if (!info->l_stream) {
			info->l_stream = g_dec_from_files ? grk_stream_create_mapped_file_read_stream(info->input_file_name)
					: grk_stream_create_mem_stream(const_cast<uint8_t*>(g_dec_in), g_dec_in_len, false, true);
			info->l_codec = info->l_stream ? grk_create_decompress(GRK_CODEC_J2K, info->l_stream) : nullptr;
			if (!info->l_codec || !grk_setup_decoder(info->l_codec, &param->core)) failed = true;
		}
		if (!failed && (info->decode_flags & GRK_DECODE_HEADER)) {
			if (!grk_read_header(info->l_codec, &info->header_info, &info->image)) failed = true;
			else if (info->init_decoders_func) return info->init_decoders_func(&info->header_info, info->image);
		}
		if (!failed && info->decode_flags != GRK_DECODE_HEADER) {
			if (info->tile) info->tile->decode_flags = info->decode_flags;
			if (!grk_set_decode_area(info->l_codec, info->image, 0, 0, 0, 0)
					|| !(grk_decode(info->l_codec, info->tile, info->image) && grk_end_decompress(info->l_codec)))
				failed = true;
		}
		if (info->decode_flags != GRK_DECODE_HEADER || failed) {
			if (info->l_stream) grk_stream_destroy(info->l_stream);
			info->l_stream = nullptr;
			if (info->l_codec) grk_destroy_codec(info->l_codec);
			info->l_codec = nullptr;
		}
		if (failed) {
			if (info->image) grk_image_destroy(info->image);
			info->image = nullptr;
			return 1;
		}
		rc = 0;
	}
	if (info->decode_flags & GRK_DECODE_POST_T1) { /* post_decode: "store" the image */
		grk_image *image = info->image;
		rc = 1;
		if (image) {
			rc = 0;
			if (g_dec_from_files && ((uint32_t) g_dec_stored >= g_dec_batch_max || image->numcomps > 3)) return 2;
			uint32_t *dims = g_dec_from_files ? g_dec_dims + 3 * g_dec_stored : g_dec_dims;
			dims[0] = image->comps[0].w; dims[1] = image->comps[0].h; dims[2] = image->numcomps;
			for (uint32_t c = 0; c < image->numcomps; ++c) {
				uint64_t n = (uint64_t) image->comps[c].w * image->comps[c].h;
				if (n > g_dec_cap || !image->comps[c].data) { rc = 2; break; }
				int32_t *dst = g_dec_from_files ? g_dec_batch_out + ((uint64_t) g_dec_stored * 3 + c) * g_dec_cap : g_dec_out[c];
				memcpy(dst, image->comps[c].data, n * sizeof(int32_t));
			}
			if (!rc) g_dec_stored = g_dec_from_files ? g_dec_stored + 1 : 1;
		}
	}
	return rc;
}

int ref_plugin_decode(const char *plugin_dir, const uint8_t *buf, uint64_t len, uint32_t reduce, uint32_t layers,
		int32_t *const *planes_out, uint64_t plane_capacity, uint32_t *dims) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	grk_plugin_load_info li;
	li.plugin_path = plugin_dir;
	if (!grk_plugin_load(li)) return -1;
	grk_plugin_init_info ii;
	ii.deviceId = 0;
	ii.verbose = true;
	if (!grk_plugin_init(ii)) { grk_plugin_cleanup(); return -1; }
	grk_decompress_parameters param;
	memset(&param, 0, sizeof(param));
	grk_set_default_decoder_parameters(&param.core);
	param.core.cp_reduce = reduce;
	param.core.cp_layer = layers;
	param.decod_format = GRK_J2K_FMT;
	param.cod_format = GRK_PXM_FMT;
	/* the plugin sizes its byte arena from the codestream FILE (as under grk_decompress -i); the decode itself reads the memory stream */
	snprintf(param.infile, sizeof(param.infile), "/tmp/grkref_plugin_%d.j2k", (int) getpid());
	{
		FILE *f = fopen(param.infile, "wb");
		if (!f || fwrite(buf, 1, len, f) != len) { if (f) fclose(f); grk_plugin_cleanup(); return -4; }
		fclose(f);
	}
	strcpy(param.outfile, "memory.ppm");
	g_dec_from_files = false;
	g_dec_in = buf; g_dec_in_len = len; g_dec_out = planes_out; g_dec_cap = plane_capacity; g_dec_dims = dims; g_dec_stored = 0;
	int32_t rc = grk_plugin_decode(&param, plugin_dec_cb);
	grk_plugin_cleanup();
	remove(param.infile);
	if (rc) return -2;
	return g_dec_stored ? 0 : -3;
}

/* ---- the official plugin path, batch decode: grk_plugin_init_batch_decode over a directory of codestreams, then
 * grk_plugin_batch_decode and polling grk_plugin_is_batch_complete like grk_decompress -y <dir> (grk_decompress.cpp:1236-1262).
 * The callback runs on the plugin's thread, file-name order; the planes of frame k land at out + (3k + c) * plane_capacity,
 * its dimensions at dims[3k..3k+2].  Returns the number of frames stored, or a negative status. */
int32_t ref_plugin_batch_decode(const char *plugin_dir, const char *in_dir, uint32_t reduce, int32_t *out, uint64_t plane_capacity,
		uint32_t *dims, uint32_t max_frames) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	const bool trace = getenv("GRK_REF_TRACE") != nullptr;
	if (trace) signal(SIGSEGV, [](int) { /* test aid: where did it crash (resolve with addr2line on the same .so files) */
		void *frames[48];
		const int n = backtrace(frames, 48);
		backtrace_symbols_fd(frames, n, 2);
		_exit(139);
	});
#define REF_TRACE(msg) do { if (trace) { fprintf(stderr, "[ref_driver] batch decode: %s\n", msg); fflush(stderr); } } while (0)
	grk_plugin_load_info li;
	li.plugin_path = plugin_dir;
	REF_TRACE("load");
	if (!grk_plugin_load(li)) return -1;
	grk_plugin_init_info ii;
	ii.deviceId = 0;
	ii.verbose = true;
	REF_TRACE("init");
	if (!grk_plugin_init(ii)) { grk_plugin_cleanup(); return -1; }
	grk_decompress_parameters param;
	memset(&param, 0, sizeof(param));
	grk_set_default_decoder_parameters(&param.core);
	param.core.cp_reduce = reduce;
	param.decod_format = GRK_J2K_FMT;
	param.cod_format = GRK_PXM_FMT;
	g_dec_from_files = true;
	g_dec_batch_out = out; g_dec_cap = plane_capacity; g_dec_dims = dims; g_dec_batch_max = max_frames; g_dec_stored = 0;
	REF_TRACE("init_batch_decode");
	int32_t rc = grk_plugin_init_batch_decode(in_dir, in_dir, &param, plugin_dec_cb);
	if (rc) { g_dec_from_files = false; grk_plugin_cleanup(); return -2; }
	REF_TRACE("batch_decode");
	grk_plugin_batch_decode(); /* (the CLI only gets here when the init call failed; the plugin tolerates both orders) */
	REF_TRACE("poll");
	while (!grk_plugin_is_batch_complete()) usleep(1000);
	REF_TRACE("stop");
	grk_plugin_stop_batch_decode();
	REF_TRACE("cleanup");
	grk_plugin_cleanup();
	REF_TRACE("done");
#undef REF_TRACE
	g_dec_from_files = false;
	return g_dec_stored;
}

/* ---- the official plugin path, batch encode: grk_plugin_batch_encode over a directory, then polling
 * grk_plugin_is_batch_complete, exactly like grk_compress -y <dir> (grk_compress.cpp:2222-2240).  The callback runs on the
 * plugin's thread, once per frame in file-name order; codestreams are appended to `out`, their lengths to lens[].
 * Returns the number of frames encoded through the plugin, or a negative status. */
static uint8_t *g_b_out = nullptr;
static uint64_t g_b_cap = 0, g_b_used = 0;
static uint64_t *g_b_lens = nullptr;
static uint32_t g_b_max = 0, g_b_count = 0, g_b_failed = 0;

static bool plugin_batch_cb(grk_plugin_encode_user_callback_info *info) {
	grk_cparameters *param = info->encoder_parameters;
	grk_image *image = info->image;
	if (!image || !info->tile || g_b_count >= g_b_max) { g_b_failed++; return false; }
	if (param->tcp_mct == 255) param->tcp_mct = (image->numcomps >= 3) ? 1 : 0;
	int64_t len = -1;
	grk_stream *stream = grk_stream_create_mem_stream(g_b_out + g_b_used, g_b_cap - g_b_used, false, false);
	grk_codec *codec = stream ? grk_create_compress(GRK_CODEC_J2K, stream) : nullptr;
	if (codec && grk_setup_encoder(codec, param, image) && grk_start_compress(codec, image)
			&& grk_encode_with_plugin(codec, info->tile) && grk_end_compress(codec))
		len = (int64_t) grk_stream_get_write_mem_stream_length(stream);
	if (stream) grk_stream_destroy(stream);
	if (codec) grk_destroy_codec(codec);
	if (len < 0) { g_b_failed++; return false; }
	g_b_lens[g_b_count++] = (uint64_t) len;
	g_b_used += (uint64_t) len;
	return true;
}

int32_t ref_plugin_batch_encode(const char *plugin_dir, const char *in_dir, uint32_t numres, uint32_t cblkw, uint32_t cblkh,
		int irreversible, uint32_t numlayers, const double *rates, uint32_t rc_algorithm, uint8_t *out, uint64_t cap,
		uint64_t *lens, uint32_t max_frames) {
	grk_set_info_handler(quiet_cb, nullptr);
	grk_set_warning_handler(quiet_cb, nullptr);
	grk_set_error_handler(quiet_cb, nullptr);
	grk_plugin_load_info li;
	li.plugin_path = plugin_dir;
	if (!grk_plugin_load(li)) return -1;
	grk_plugin_init_info ii;
	ii.deviceId = 0;
	ii.verbose = true;
	if (!grk_plugin_init(ii)) { grk_plugin_cleanup(); return -1; }
	grk_cparameters param;
	grk_set_default_encoder_parameters(&param);
	param.decod_format = GRK_PXM_FMT;
	param.cod_format = GRK_J2K_FMT;
	param.numresolution = numres;
	param.cblockw_init = cblkw;
	param.cblockh_init = cblkh;
	param.irreversible = irreversible != 0;
	param.rateControlAlgorithm = rc_algorithm;
	param.tcp_numlayers = numlayers ? numlayers : 1;
	for (uint32_t i = 0; i < numlayers; ++i) param.tcp_rates[i] = rates[i];
	if (!numlayers) param.tcp_rates[0] = 0;
	param.cp_disto_alloc = 1;
	param.tcp_mct = 255;
	g_b_out = out; g_b_cap = cap; g_b_used = 0; g_b_lens = lens; g_b_max = max_frames; g_b_count = 0; g_b_failed = 0;
	int32_t rc = grk_plugin_batch_encode(in_dir, in_dir, &param, plugin_batch_cb);
	if (rc) { grk_plugin_cleanup(); return -2; }
	while (!grk_plugin_is_batch_complete()) usleep(1000);
	grk_plugin_stop_batch_encode();
	grk_plugin_cleanup();
	return g_b_failed ? -3 : (int32_t) g_b_count;
}

} /* extern "C" */
