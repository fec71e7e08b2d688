/* Hand-written stand-in for the header the reference's cmake step would generate
 * (src/lib/jp2/grk_config.h.cmake.in).  Only version strings and SIMD-probe results. */
#pragma once
#define GROK_HAVE_STDINT_H 1
#define GRK_VERSION_MAJOR 5
#define GRK_VERSION_MINOR 1
#define GRK_VERSION_BUILD 0
#define GROK_PLUGIN_NAME "grok_plugin"
#define AVX2_FOUND "true"
#define AVX_FOUND "true"
#define SSE4_1_FOUND "true"
#define SSE3_FOUND "true"
