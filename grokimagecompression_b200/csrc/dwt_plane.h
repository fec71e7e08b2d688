// The DWT kernels' view of one tile-component plane at one decomposition level (shared with the CPU emulation
// harness tests/dwt_emu.cpp, hence its own header).
#pragma once
#include <cstdint>

namespace gb {

// One tile-component plane of one decomposition level, as the DWT kernels see it.
// (a-4..a-7 of SURVEY.md section 8: WaveletForward.h:40-161, dwt.cpp:724-858, 1544-1738)
struct DwtPlane {
	const int32_t *src;   // forward: samples of the level-l LL region; inverse: LL_{l+1} (low-low quadrant)
	const int32_t *band;  // inverse only: buffer holding HL/LH/HH of this level in Mallat position
	int32_t *dst;         // forward: Mallat layout of this level; inverse: reconstructed LL_l
	uint32_t src_stride, band_stride, dst_stride;
	uint32_t rw, rh;      // size of the level-l region
	uint32_t sw, sh;      // low-pass counts (size of LL_{l+1})
	uint32_t cas_x, cas_y; // parity of the region origin on the canvas (1: first sample is high-pass)
	uint32_t tiles_x, tiles_y; // CTA tiling of this plane
	uint32_t first_cta;   // prefix sum of CTAs over the planes of the launch
	uint32_t pad;
};

} // namespace gb
