// Internal declarations shared by the kernels and the C-ABI layer of libgrok_b200.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "dwt_plane.h"

namespace gb {

// Encoder-side code block (a-8, a-9 of SURVEY.md section 8: T1Part1.cpp:58-133, t1.cpp:1182-1326)
struct EncBlock {
	const int32_t *src;  // top-left coefficient of the block inside its sub-band (DWT output)
	uint32_t stride;
	uint16_t w, h;
	uint8_t orient, reversible, sty /* code-block style switches, STY_* */, band_numbps /* band->numbps: the HT coder's missing MSBs */;
	int32_t inv_step;
	uint32_t pass_offset; // first slot in rates/dists
	uint32_t max_passes;
	uint32_t scratch_cap; // byte capacity reserved for this block
	uint64_t scratch_off; // byte offset of the block's scratch area (one pad byte precedes the stream)
	double rd_weight;
	uint64_t sym_off;     // byte offset of the block's (context, decision) stream, 16-byte aligned
	uint32_t sym_cap;     // bytes reserved for it (t1_symbol_capacity)
	float stepsize;       // band->stepsize: the HT path quantises with 1 / stepsize in float (T1HT.cpp:88-92)
};

// code-block style switches (tccp->cblk_sty, grok.h GRK_CBLKSTY_*)
enum { STY_LAZY = 1, STY_RESET = 2, STY_TERMALL = 4, STY_VSC = 8, STY_PTERM = 16, STY_SEGSYM = 32, STY_ALL = 63, STY_HT = 64 /* HTJ2K block coder: not combinable with the others */ };

struct EncResult { // == gb200_cblk_enc
	uint32_t numbps, numpasses, data_len, decisions;
	uint64_t data_offset;
};

// Decoder-side code block (a-13: t1.cpp:1038-1130, T1Part1.cpp:135-329)
struct DecBlock {
	int32_t *dst;        // top-left of the block inside the coefficient plane
	uint32_t stride;
	uint16_t w, h;
	uint8_t orient, reversible, sty, roishift /* ROI up-shift of the component (decoding starts roishift planes higher) */;
	float stepsize;
	uint32_t band_numbps; // band->numbps: the HT decoder's missing MSBs are band_numbps - numbps of the block
};

struct DecSeg { // == gb200_cblk_seg: one codeword segment of a block (TERMALL / LAZY streams)
	uint32_t len, numpasses;
};

struct DecInput { // == gb200_cblk_dec
	uint32_t numbps, numpasses, data_len, reserved;
	uint64_t data_offset;
};

// tables.cu
void upload_tables();

// mct.cu : DC level shift fused with RCT/ICT.  n samples per plane.
void launch_dcshift_fwd(int32_t *x, uint64_t n, int32_t shift, int reversible, cudaStream_t s);
void launch_mct_fwd(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n, int32_t s0, int32_t s1, int32_t s2,
		int reversible, int do_shift, cudaStream_t s);
void launch_dcshift_inv(int32_t *x, uint64_t n, int32_t shift, int reversible, int32_t lo, int32_t hi, cudaStream_t s);
void launch_mct_inv(int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n, const int32_t shift[3], const int32_t lo[3],
		const int32_t hi[3], int reversible, int do_shift_clamp, cudaStream_t s);

// narrow-sample boundary: packed 8 / 16-bit samples (sample_bytes 1 or 2, sgnd: int8 / int16) widened on load, narrowed on store
void launch_mct_fwd_packed(const void *s0, const void *s1, const void *s2, int32_t *c0, int32_t *c1, int32_t *c2, uint64_t n,
		int32_t sh0, int32_t sh1, int32_t sh2, int reversible, uint32_t sample_bytes, int sgnd, cudaStream_t s);
void launch_mct_inv_packed(const int32_t *c0, const int32_t *c1, const int32_t *c2, void *d0, void *d1, void *d2, uint64_t n,
		const int32_t shift[3], const int32_t lo[3], const int32_t hi[3], int reversible, uint32_t sample_bytes, int sgnd, cudaStream_t s);
void launch_dcshift_fwd_packed(const void *src, int32_t *dst, uint64_t n, int32_t shift, int reversible, uint32_t sample_bytes, int sgnd,
		cudaStream_t s);
void launch_dcshift_inv_packed(const int32_t *src, void *dst, uint64_t n, int32_t shift, int reversible, int32_t lo, int32_t hi,
		uint32_t sample_bytes, int sgnd, cudaStream_t s);

// dwt.cu : one launch = one decomposition level of every plane in `planes` (device array).  Streaming kernels
// (dwt_stream.cuh): one warp per work item of `rows` rows by dwt_stream_shape() columns, `unroll` row pairs prefetched
// (1, 2 or 4), `halo_lanes` 1 or 2; total_items counts work items.  DwtPlane::tiles_x / tiles_y must have been computed for
// the same shape (over rw + cas_x / rh + cas_y).
void launch_dwt_fwd(const DwtPlane *planes_dev, const uint32_t *item_plane_dev, uint32_t total_items, int reversible, int rows,
		int unroll, int halo_lanes, cudaStream_t s);
void launch_dwt_inv(const DwtPlane *planes_dev, const uint32_t *item_plane_dev, uint32_t total_items, int reversible, int rows,
		int unroll, int halo_lanes, cudaStream_t s);
int dwt_stream_warps_per_sm(int reversible, int forward, int unroll); // work items resident per SM (occupancy of that kernel)
void dwt_stream_shape(int halo_lanes, uint32_t *tw); // valid columns per work item of the streaming kernels (halo_lanes 1 or 2)

// t1_enc.cu / t1_dec.cu
uint32_t t1_symbol_capacity(uint32_t w, uint32_t h, uint32_t planes);
// styles: non-zero when any block of the table has a code-block style switch set (selects the general MQ kernel)
void launch_t1_encode(const EncBlock *blocks, uint32_t nblocks, int rate_control, int styles, uint8_t *symbols, uint8_t *scratch,
		EncResult *results, uint32_t *rates, double *dists, cudaStream_t s);
// total (may be NULL): receives the byte count of the compacted data
void launch_t1_gather(const EncBlock *blocks, EncResult *results, uint32_t nblocks, const uint8_t *scratch,
		uint8_t *data, uint64_t *total, cudaStream_t s);
// ht.cu : the HTJ2K block coder (cleanup pass), one thread per block; the encoder leaves its bytes in the block's scratch area
// like the MQ coder does (launch_t1_gather compacts them); the decoder writes de-quantised samples (no clear / finish pass)
uint32_t t1_ht_scratch_extra(); // bytes an HT block needs beyond 4 w h: the MEL and VLC staging areas
void launch_t1_ht_encode(const EncBlock *blocks, uint32_t nblocks, uint8_t *scratch, EncResult *results, uint32_t *rates, double *dists,
		cudaStream_t s);
void launch_t1_ht_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data, cudaStream_t s);
// rd.cu : feasible truncation points and 8.8 log slopes of every block (RateControl.cpp:31-168); cache: one double per pass slot
void launch_rd_slopes(const EncBlock *blocks, const EncResult *results, uint32_t nblocks, const uint32_t *rates, const double *dists,
		uint16_t *slopes, double *cache, cudaStream_t s);
// three launches (clear, decode, de-quantise); max_w / max_h: largest block of the table; `data` must stay
// readable for T1_DEC_DATA_SLACK bytes past the last segment.  Returns non-zero if a block cannot be placed.
constexpr int T1_DEC_LAUNCHES = 3;
constexpr size_t T1_DEC_DATA_SLACK = 64;
// styles: some block has a style switch set; seg_start (nblocks + 1 prefix offsets) / segs: codeword segments, or NULL when
// every block is a single segment
int launch_t1_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data,
		uint32_t max_w, uint32_t max_h, int styles, const uint32_t *seg_start, const DecSeg *segs, cudaStream_t s);

} // namespace gb
