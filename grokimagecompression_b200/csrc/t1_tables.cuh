// Tier-1 tables, regenerated at start-up from the defining rules of ISO/IEC 15444-1
// (Table C.2, Tables D.1-D.4) and the reference's distortion-estimate formulas
// (t1_generate_luts.cpp:63-209, 291-317); nothing is copied from t1_luts.h.
// Each translation unit that includes this header owns a private __constant__ copy.
#pragma once
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <mutex>

namespace gb {

enum { CTX_ZC0 = 0, CTX_SC0 = 9, CTX_MR0 = 14, CTX_AGG = 17, CTX_UNI = 18, NCTX = 19 };

// MQ probability state machine: qe | nmps << 16 | nlps << 22 | switch << 28   (Table C.2)
static __constant__ uint32_t c_mq[47];
// zero-coding context by orientation and 8 neighbour bits: NW N NE W E SW S SE = bits 0..7
static __constant__ uint8_t c_zc[4][256];
// sign coding: index = sigN sigW sigE sigS negN negW negE negS (bits 0..7); value = ctx | xorbit << 5
static __constant__ uint8_t c_sc[256];
// distortion estimates: [0] sig (plane>0) [1] sig (plane 0) [2] ref (plane>0) [3] ref (plane 0)
static __constant__ int16_t c_nmsedec[4][128];

static inline void build_and_upload_t1_tables() {
	static const uint16_t qe[47] = {0x5601, 0x3401, 0x1801, 0x0AC1, 0x0521, 0x0221, 0x5601, 0x5401, 0x4801, 0x3801,
		0x3001, 0x2401, 0x1C01, 0x1601, 0x5601, 0x5401, 0x5101, 0x4801, 0x3801, 0x3401, 0x3001, 0x2801, 0x2401, 0x2201,
		0x1C01, 0x1801, 0x1601, 0x1401, 0x1201, 0x1101, 0x0AC1, 0x09C1, 0x08A1, 0x0521, 0x0441, 0x02A1, 0x0221, 0x0141,
		0x0111, 0x0085, 0x0049, 0x0025, 0x0015, 0x0009, 0x0005, 0x0001, 0x5601};
	static const uint8_t nmps[47] = {1, 2, 3, 4, 5, 38, 7, 8, 9, 10, 11, 12, 13, 29, 15, 16, 17, 18, 19, 20, 21, 22, 23,
		24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 45, 46};
	static const uint8_t nlps[47] = {1, 6, 9, 12, 29, 33, 6, 14, 14, 14, 17, 18, 20, 21, 14, 14, 15, 16, 17, 18, 19, 19,
		20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 46};
	uint32_t mq[47];
	for (int i = 0; i < 47; ++i)
		mq[i] = qe[i] | ((uint32_t) nmps[i] << 16) | ((uint32_t) nlps[i] << 22) | ((i == 0 || i == 6 || i == 14) ? 1u << 28 : 0);
	cudaMemcpyToSymbol(c_mq, mq, sizeof(mq));

	uint8_t zc[4][256];
	for (int orient = 0; orient < 4; ++orient)
		for (int f = 0; f < 256; ++f) {
			int h = ((f >> 3) & 1) + ((f >> 4) & 1);
			int v = ((f >> 1) & 1) + ((f >> 6) & 1);
			int d = (f & 1) + ((f >> 2) & 1) + ((f >> 5) & 1) + ((f >> 7) & 1);
			if (orient == 1) { int t = h; h = v; v = t; } // HL: roles of horizontal and vertical swap
			int n;
			if (orient == 3) {
				int hv = h + v;
				if (d == 0) n = hv >= 2 ? 2 : hv;
				else if (d == 1) n = hv >= 2 ? 5 : 3 + hv;
				else if (d == 2) n = hv >= 1 ? 7 : 6;
				else n = 8;
			} else {
				if (h == 2) n = 8;
				else if (h == 1) n = v >= 1 ? 7 : (d >= 1 ? 6 : 5);
				else if (v == 2) n = 4;
				else if (v == 1) n = 3;
				else n = d >= 2 ? 2 : d;
			}
			zc[orient][f] = (uint8_t) n;
		}
	cudaMemcpyToSymbol(c_zc, zc, sizeof(zc));

	uint8_t sc[256];
	for (int f = 0; f < 256; ++f) {
		auto contrib = [&](int sigbit, int negbit) { return ((f >> sigbit) & 1) ? (((f >> negbit) & 1) ? -1 : 1) : 0; };
		int hc = contrib(1, 5) + contrib(2, 6), vc = contrib(0, 4) + contrib(3, 7);
		hc = hc > 1 ? 1 : (hc < -1 ? -1 : hc);
		vc = vc > 1 ? 1 : (vc < -1 ? -1 : vc);
		int x = (hc < 0 || (hc == 0 && vc < 0)) ? 1 : 0;
		if (hc < 0) { hc = -hc; vc = -vc; }
		int ctx = hc == 0 ? CTX_SC0 + (vc == 0 ? 0 : 1) : CTX_SC0 + 3 + vc;
		sc[f] = (uint8_t) (ctx | (x << 5));
	}
	cudaMemcpyToSymbol(c_sc, sc, sizeof(sc));

	int16_t nm[4][128];
	for (int i = 0; i < 128; ++i) {
		double t = i / 64.0;
		auto fix = [](double e) { int r = (int) (std::floor(e * 64.0 + 0.5) / 64.0 * 8192.0); return (int16_t) (r < 0 ? 0 : r); };
		double u = t, v = t - 1.5;
		nm[0][i] = fix(u * u - v * v);
		nm[1][i] = fix(u * u);
		u = t - 1.0;
		v = (i & 64) ? t - 1.5 : t - 0.5;
		nm[2][i] = fix(u * u - v * v);
		nm[3][i] = fix(u * u);
	}
	cudaMemcpyToSymbol(c_nmsedec, nm, sizeof(nm));
}

// __constant__ memory is per device and a process may hold contexts on several (gb200_create(device)): the tables of this
// translation unit are uploaded once per device, under a lock (contexts are driven from several host threads)
static inline void ensure_t1_tables() {
	static std::mutex mu;
	static uint64_t ready[4] = {0, 0, 0, 0}; // bit per device ordinal
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(mu);
	if (dev >= 0 && dev < 256 && (ready[dev >> 6] >> (dev & 63) & 1)) return;
	build_and_upload_t1_tables();
	cudaDeviceSynchronize(); // cudaMemcpyToSymbol from pageable memory has completed on return; be explicit
	if (dev >= 0 && dev < 256) ready[dev >> 6] |= 1ull << (dev & 63);
}

} // namespace gb
