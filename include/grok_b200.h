/*
 * grok_b200.h -- C ABI of libgrok_b200.so: Grok's tile-coding hot path on one B200.
 *
 * The library replaces, for one tile or a batch of tiles/frames, the stage calls that Grok's tile
 * coder/decoder makes (all paths relative to /root/reference/src/lib/jp2):
 *
 *   encode  TileProcessor::encode_tile  TileProcessor.cpp:994-1012
 *             dc_level_shift_encode     TileProcessor.cpp:1449-1471
 *             mct_encode                TileProcessor.cpp:1473-1518 -> mct.cpp:85,195
 *             dwt_encode                TileProcessor.cpp:1520-1533 -> WaveletForward.h:40
 *             t1_encode                 TileProcessor.cpp:1535-1557 -> Tier1.cpp:24, T1Part1.cpp:58,96, t1.cpp:1182
 *   decode  TileProcessor::decode_tile  TileProcessor.cpp:1141-1177
 *             Tier1::decodeCodeblocks   Tier1.cpp:177 -> T1Part1.cpp:135,199, t1.cpp:1038
 *             Wavelet::decode           dwt.cpp:1208 (5/3), 2154 (9/7)
 *             mct_decode                TileProcessor.cpp:1303-1375 -> mct.cpp:143,352
 *             dc_level_shift_decode     TileProcessor.cpp:1377-1432
 *
 * Everything is plain C: pointers and sizes only.  Quantisation parameters (step sizes, inverse
 * steps, band bit depths, R/D weights) are INPUTS computed by the host codec
 * (Quantizer.cpp:65-105); Tier-2, PCRD and codestream I/O stay in the host.
 * All functions return 0 on success; on failure a non-zero code, and gb200_last_error() describes
 * it.  There is no CPU fallback: without a usable CUDA device gb200_create() fails.
 */
#ifndef GROK_B200_H
#define GROK_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_MAX_RES 33
#define GB200_MAX_BANDS (3 * GB200_MAX_RES - 2)
#define GB200_ABI_VERSION 3
#if defined(__GNUC__)
#define GB200_API __attribute__((visibility("default")))
#else
#define GB200_API
#endif

enum {
	GB200_OK = 0,
	GB200_ERR_CUDA = 1,        /* CUDA runtime error (text in gb200_last_error) */
	GB200_ERR_PARAM = 2,       /* invalid argument */
	GB200_ERR_UNSUPPORTED = 3, /* code-block style / geometry outside this build's scope */
	GB200_ERR_CAPACITY = 4,    /* caller's output buffer too small */
	GB200_ERR_NOMEM = 5
};

typedef struct gb200_ctx gb200_ctx;   /* one CUDA device + stream + constant tables */
typedef struct gb200_plan gb200_plan; /* geometry, block table and device buffers of a tile batch */

/* One tile-component: what TileComponent::init (TileComponent.cpp:165-507) derives from the
 * codestream parameters, plus the per-band quantisation constants. Band order is the host's:
 * index 0 = LL of resolution 0, then for each resolution r>=1: HL, LH, HH (3r-2, 3r-1, 3r). */
typedef struct gb200_comp_params {
	uint32_t x0, y0, x1, y1;            /* tile-component rectangle on the component canvas */
	uint32_t numres;                    /* tccp->numresolutions (decompositions + 1) */
	uint32_t cblkw_expn, cblkh_expn;    /* tccp->cblkw / cblkh (log2 nominal code-block size) */
	uint32_t prcw_expn[GB200_MAX_RES];  /* tccp->prcw[resno] */
	uint32_t prch_expn[GB200_MAX_RES];  /* tccp->prch[resno] */
	uint32_t qmfbid;                    /* 1 = reversible 5/3, 0 = irreversible 9/7 */
	uint32_t prec;                      /* image component precision */
	uint32_t sgnd;                      /* image component signedness */
	int32_t dc_shift;                   /* tccp->m_dc_level_shift */
	uint32_t cblk_sty;                  /* tccp->cblk_sty: LAZY 0x01, RESET 0x02, TERMALL 0x04, VSC 0x08, PTERM 0x10, SEGSYM 0x20
	                                     * (t1.cpp:1131-1151, 1223-1298); HT 0x40 alone = the HTJ2K block coder (t1/t1_ht, every
	                                     * component of a plan or none): the encoder emits the cleanup pass as one pass with numbps 1
	                                     * like T1HT::encode, the decoder takes single-pass blocks; band_numbps and stepsize must
	                                     * then be the host's values on BOTH sides (k_msbs = band_numbps - numbps, Tier1.cpp:86,166) */
	uint32_t roishift;                  /* tccp->roishift (max-shift ROI): the encoder only sees it through band_numbps, the decoder
	                                     * starts roishift planes higher and shifts the ROI samples back (T1Part1.cpp:230-252) */
	float stepsize[GB200_MAX_BANDS];    /* band->stepsize (decoder side already carries the x0.5) */
	uint32_t inv_step[GB200_MAX_BANDS]; /* band->inv_step, 13-bit fixed point */
	uint32_t band_numbps[GB200_MAX_BANDS]; /* band->numbps (upper bound of a block's bit planes) */
	double rd_weight[GB200_MAX_BANDS];  /* (mct_norm * dwt_norm) * stepsize of t1_getwmsedec, t1.cpp:912-932 */
} gb200_comp_params;

typedef struct gb200_tile_params {
	uint32_t numcomps;
	uint32_t mct;           /* tcp->mct: 0 none, 1 = RCT (5/3) or ICT (9/7) on components 0..2 */
	uint32_t rate_control;  /* TileProcessor::needs_rate_control(): fill per-pass distortion */
	uint32_t numres_decode; /* decoder: resolutions to reconstruct (tilec->minimum_num_resolutions); 0 = all */
	const gb200_comp_params *comps; /* numcomps entries */
} gb200_tile_params;

/* Static description of one code block, in the host's traversal order
 * (tile, comp, resno, band, precinct, block: plugin_bridge.cpp:148-152, Tier1.cpp:39-91). */
typedef struct gb200_cblk_info {
	uint32_t tileno, compno, resno, bandno /* orient 0..3 */, precno, cblkno;
	uint32_t x0, y0, x1, y1; /* band coordinates (grk_tcd_cblk_enc::x0..y1) */
	uint32_t band_index;     /* index into gb200_comp_params::stepsize etc. */
	uint32_t pass_offset;    /* first slot of this block in the rates[] / dists[] arrays */
	uint32_t max_passes;     /* slots reserved: 3 * band_numbps - 2 */
} gb200_cblk_info;

/* Encoder result for one code block (grk_tcd_cblk_enc after T1Part1::encode, T1Part1.cpp:96-133) */
typedef struct gb200_cblk_enc {
	uint32_t numbps;       /* cblk->numbps */
	uint32_t numpasses;    /* cblk->num_passes_encoded */
	uint32_t data_len;     /* bytes of the MQ segment == rate of the last pass */
	uint32_t decisions;    /* MQ decisions coded (work counter; not part of the reference contract) */
	uint64_t data_offset;  /* offset of the block's bytes in the data buffer */
} gb200_cblk_enc;

/* Decoder input for one code block (what T2 parsed: grk_tcd_cblk_dec, T1Part1.cpp:135-197) */
typedef struct gb200_cblk_dec {
	uint32_t numbps;      /* cblk->numbps (roishift already removed) */
	uint32_t numpasses;   /* passes to decode (single MQ segment) */
	uint32_t data_len;    /* concatenated segment bytes */
	uint32_t reserved;
	uint64_t data_offset; /* offset in the data buffer */
} gb200_cblk_dec;

/* One codeword segment of a code block (grk_tcd_seg: len, numpasses).  Blocks coded with TERMALL or LAZY arrive from
 * Tier-2 as several segments (T2.cpp:835-851); their bytes are concatenated at gb200_cblk_dec::data_offset. */
typedef struct gb200_cblk_seg {
	uint32_t len;
	uint32_t numpasses;
} gb200_cblk_seg;

/* ---- context ---------------------------------------------------------------------------------- */
GB200_API int gb200_abi_version(void);
GB200_API const char *gb200_last_error(void); /* thread-local, never NULL */
/* number of usable CUDA devices (0 when there is none or the driver is missing).  A context belongs to ONE device; a host
 * that wants several GPUs creates one context per device and drives each from its own thread, dealing independent units
 * (tiles, frames) to them -- there is no cross-device state in this library ("Several GPUs"). */
GB200_API int gb200_device_count(void);
GB200_API int gb200_create(int device, gb200_ctx **out);
GB200_API void gb200_destroy(gb200_ctx *ctx);
/* number of kernels this context has launched so far (for bench.py's gpu_launches) */
GB200_API uint64_t gb200_launch_count(const gb200_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is issued on */
GB200_API void *gb200_stream(const gb200_ctx *ctx);

/* ---- plans ------------------------------------------------------------------------------------ */
GB200_API int gb200_plan_create(gb200_ctx *ctx, uint32_t ntiles, const gb200_tile_params *tiles, int is_encoder,
		gb200_plan **out);
GB200_API void gb200_plan_destroy(gb200_plan *plan);
GB200_API uint64_t gb200_plan_num_blocks(const gb200_plan *plan);
GB200_API uint64_t gb200_plan_num_pass_slots(const gb200_plan *plan);
GB200_API uint64_t gb200_plan_num_samples(const gb200_plan *plan); /* sum of tile-component areas */
GB200_API const gb200_cblk_info *gb200_plan_blocks(const gb200_plan *plan);
GB200_API uint64_t gb200_plan_data_capacity(const gb200_plan *plan); /* worst-case encoder output bytes */
/* the code blocks of one tile-component in the host's traversal order (resno, band, precinct, block), the geometry
 * gb200_plan_create builds its block table from (TileComponent.cpp:193-489): pure host code, needs no device.  Fills
 * out[0 .. min(count, cap)) (tileno, compno = 0; pass_offset counted from 0) and returns the count; numres_limit = the
 * resolutions to enumerate (0 = all). */
GB200_API uint64_t gb200_enumerate_blocks(const gb200_comp_params *comp, uint32_t numres_limit, gb200_cblk_info *out, uint64_t cap);
/* precinct grid of one resolution of a tile-component (grk_tcd_resolution::pw, ph: TileComponent.cpp:303-328), the
 * number of precincts every band of that resolution has in the host's tree, empty ones included.  Pure geometry: needs
 * no device. */
GB200_API int gb200_precinct_grid(const gb200_comp_params *comp, uint32_t resno, uint32_t *pw, uint32_t *ph);

/* ---- whole path, host buffers (the call the host TCD makes per tile batch) ------------------- */
/* planes[t * numcomps + c]: int32 samples of tile t / component c, row stride = x1 - x0
 * (grk_image_comp::data as copied into the tile buffer, j2k.cpp:2077-2108). */
GB200_API int gb200_encode_tiles(gb200_plan *plan, const int32_t *const *planes, gb200_cblk_enc *blocks, uint32_t *rates,
		double *dists, uint8_t *data, uint64_t data_capacity, uint64_t *data_len);
/* planes_out[t * numcomps + c]: decoded int32 samples, row stride = width of the decoded resolution */
GB200_API int gb200_decode_tiles(gb200_plan *plan, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len,
		int32_t *const *planes_out);

/* ---- narrow-sample boundary --------------------------------------------------------------------
 * The reference expands every image sample to int32 on the host before its tile coder sees it and narrows the decoded
 * tile back afterwards (TileProcessor.cpp:1201-1258 copy-in, 1691-1921 copy-out); over PCIe that is 4 bytes per 8-bit
 * sample in each direction.  A plan set to 1 or 2 bytes per sample takes / returns the packed samples themselves
 * (uint8 / uint16, int8 / int16 for signed components; row stride = width) and widens / clamps + narrows them inside the
 * level-shift + MCT pass on the device.  Every component's precision must fit the type.  4 (the default) = int32 planes. */
GB200_API int gb200_plan_set_sample_bytes(gb200_plan *plan, uint32_t sample_bytes);
GB200_API int gb200_encode_tiles_packed(gb200_plan *plan, const void *const *planes, gb200_cblk_enc *blocks, uint32_t *rates,
		double *dists, uint8_t *data, uint64_t data_capacity, uint64_t *data_len);
GB200_API int gb200_decode_tiles_packed(gb200_plan *plan, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len,
		void *const *planes_out);
GB200_API int gb200_encode_upload_packed(gb200_plan *plan, const void *const *planes);
GB200_API int gb200_decode_download_packed(gb200_plan *plan, void *const *planes_out);

/* ---- the same path split in upload / run / download, so a caller can time the device part ---- */
GB200_API int gb200_encode_upload(gb200_plan *plan, const int32_t *const *planes);
GB200_API int gb200_encode_run(gb200_plan *plan);      /* asynchronous on gb200_stream() */
GB200_API int gb200_encode_download(gb200_plan *plan, gb200_cblk_enc *blocks, uint32_t *rates, double *dists, uint8_t *data,
		uint64_t data_capacity, uint64_t *data_len);
/* PCRD preparation on the device, replaces the per-block RateControl::convexHull calls of the rate allocator
 * (t2/RateControl.cpp:31-118, called from TileProcessor.cpp:409): for every code block of the last encode run, which
 * passes are feasible truncation points and their distortion-rate slopes as ln(slope) in 8.8 fixed point (slopeToLog,
 * RateControl.cpp:159-168); 0 = not a truncation point.  slopes has gb200_plan_num_pass_slots() entries, pass p of block i at
 * gb200_cblk_info::pass_offset + p (= grk_tcd_pass::slope).  Needs rates and distortions, i.e. a plan with rate control on. */
GB200_API int gb200_encode_slopes(gb200_plan *plan, uint16_t *slopes);
GB200_API int gb200_decode_upload(gb200_plan *plan, const gb200_cblk_dec *blocks, const uint8_t *data, uint64_t data_len);
/* codeword segments for the next gb200_decode_upload / gb200_decode_tiles: block i owns segs[seg_start[i] .. seg_start[i+1])
 * (seg_start has num_blocks + 1 entries); gb200_cblk_dec::numpasses / data_len stay the totals.  NULL, NULL = every
 * block is one segment (the default). */
GB200_API int gb200_decode_set_segments(gb200_plan *plan, const uint32_t *seg_start, const gb200_cblk_seg *segs);
GB200_API int gb200_decode_run(gb200_plan *plan);      /* asynchronous on gb200_stream() */
GB200_API int gb200_decode_download(gb200_plan *plan, int32_t *const *planes_out);
GB200_API int gb200_sync(gb200_ctx *ctx);
/* keep / bring back a device-side copy of the uploaded input planes, so that gb200_encode_run can be
 * repeated on HBM-resident input (the run transforms the planes in place) */
GB200_API int gb200_encode_stash(gb200_plan *plan);
GB200_API int gb200_encode_restore(gb200_plan *plan); /* asynchronous device-to-device copy */
/* run only one stage of an uploaded plan (for per-kernel timing): 0 dc+mct, 1 dwt, 2 t1 */
GB200_API int gb200_encode_run_stage(gb200_plan *plan, int stage);
GB200_API int gb200_decode_run_stage(gb200_plan *plan, int stage);
/* after gb200_encode_run: copy the transformed (Mallat-layout) coefficient plane of tile t / comp c
 * to host (row stride = x1-x0) -- the tile buffer the reference holds after dwt_encode() */
GB200_API int gb200_encode_get_coefficients(gb200_plan *plan, uint32_t tileno, uint32_t compno, int32_t *out);
/* before gb200_decode_run_stage(1): overwrite the coefficient plane (what T1 decode would have left) */
GB200_API int gb200_decode_set_coefficients(gb200_plan *plan, uint32_t tileno, uint32_t compno, const int32_t *in);

/* ---- stage-level entry points on host buffers, in place: one per reference stage function ---- */
/* grk::mct::encode_rev / decode_rev / encode_irrev / decode_irrev   mct.cpp:85,143,195,352 */
GB200_API int gb200_mct_encode_rev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n);
GB200_API int gb200_mct_decode_rev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n);
GB200_API int gb200_mct_encode_irrev(gb200_ctx *ctx, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n);
GB200_API int gb200_mct_decode_irrev(gb200_ctx *ctx, float *c0, float *c1, float *c2, uint64_t n);
/* TileProcessor::dc_level_shift_encode / _decode for one component   TileProcessor.cpp:1449,1377 */
GB200_API int gb200_dc_shift_encode(gb200_ctx *ctx, int32_t *x, uint64_t n, int32_t shift, int qmfbid);
GB200_API int gb200_dc_shift_decode(gb200_ctx *ctx, int32_t *x, uint64_t n, int32_t shift, int qmfbid, int32_t lo, int32_t hi);
/* grk::Wavelet::encode / decode for one tile-component   Wavelet.cpp:35, dwt.cpp:1208,2154 */
GB200_API int gb200_dwt_encode(gb200_ctx *ctx, int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
		uint32_t numres, int qmfbid);
GB200_API int gb200_dwt_decode(gb200_ctx *ctx, int32_t *buf, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
		uint32_t numres, uint32_t numres_decode, int qmfbid);

/* Tier-1 on a list of independent blocks taken from one int32 plane (T1Part1::preEncode + encode). */
typedef struct gb200_t1_block {
	uint32_t x, y, w, h;  /* position and size inside the plane */
	uint32_t orient;      /* 0 LL, 1 HL, 2 LH, 3 HH */
	uint32_t qmfbid;      /* 1: coefficient * 64;  0: fixed-point multiply by inv_step */
	uint32_t inv_step;
	float stepsize;       /* decoder: de-quantisation step (9/7) */
	double rd_weight;
	uint32_t cblk_sty;    /* code-block style switches, as gb200_comp_params::cblk_sty */
	uint32_t roishift;    /* decoder: ROI up-shift of the component */
} gb200_t1_block;
GB200_API int gb200_t1_encode_blocks(gb200_ctx *ctx, const int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, int rate_control, uint32_t max_passes, gb200_cblk_enc *results,
		uint32_t *rates, double *dists, uint8_t *data, uint64_t data_capacity, uint64_t *data_len);
/* T1Part1::decode + postDecode into a zeroed int32 plane */
GB200_API int gb200_t1_decode_blocks(gb200_ctx *ctx, int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, const gb200_cblk_dec *inputs, const uint8_t *data, uint64_t data_len);
/* the same with codeword segments (seg_start: nblocks + 1 prefix offsets into segs; NULL, NULL = single segments) */
GB200_API int gb200_t1_decode_blocks_segs(gb200_ctx *ctx, int32_t *plane, uint32_t width, uint32_t height, uint32_t nblocks,
		const gb200_t1_block *blocks, const gb200_cblk_dec *inputs, const uint32_t *seg_start, const gb200_cblk_seg *segs,
		const uint8_t *data, uint64_t data_len);

#ifdef __cplusplus
}
#endif
#endif /* GROK_B200_H */
