/*
 * ref_tap.cpp -- TEST / BENCH INFRASTRUCTURE.  A tap on the UNMODIFIED reference's tile coder (oracle/_ref/libgrok_ref.so):
 * loaded in front of it (dlopen RTLD_GLOBAL before the driver, or LD_PRELOAD), it defines the stage functions that
 * TileProcessor::encode_tile / decode_tile call (TileProcessor.cpp:994-1012, 1141-1177), forwards each call to the reference's
 * own definition (dlsym RTLD_NEXT on the mangled name) and records
 *   - the wall clock spent inside each stage: the reference's HOT PATH ONLY (level shift, MCT, DWT, Tier-1 and their
 *     inverses), i.e. exactly the work libgrok_b200.so replaces, next to the whole encode_tile / decode_tile time, which also
 *     holds the host-side PCRD / Tier-2 ("the B2 interposition tap");
 *   - whether the block table the product derives from the coding parameters (gb200_enumerate_blocks, pure host code)
 *     equals the reference's own tile structure, block by block -- a geometry parity check that needs no GPU.
 * Nothing here computes any result: every stage is the reference's.
 */
#include "grok_includes.h"
#include "Tier1.h"
#include "T1Interface.h"
#include "../include/grok_b200.h"
#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <vector>

namespace {

enum { S_DC_ENC, S_MCT_ENC, S_DWT_ENC, S_T1_ENC, S_T1_DEC, S_DWT_DEC, S_MCT_DEC, S_DC_DEC, S_ENCODE_TILE, S_DECODE_TILE,
	S_SIMULATE /* T2::encode_packets_simulate: the serial part of the rate allocation, inside S_ENCODE_TILE */, S_COUNT };
double g_secs[S_COUNT] = {0};
uint64_t g_calls[S_COUNT] = {0};
uint64_t g_geo_blocks = 0, g_geo_mismatch = 0, g_geo_tilecomps = 0;
std::mutex g_mu;

struct Timer {
	int i;
	std::chrono::steady_clock::time_point t0;
	explicit Timer(int idx) : i(idx), t0(std::chrono::steady_clock::now()) {}
	~Timer() {
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		std::lock_guard<std::mutex> lk(g_mu);
		g_secs[i] += dt;
		g_calls[i]++;
	}
};

template<typename Fn> Fn next(const char *mangled) {
	Fn f = (Fn) dlsym(RTLD_NEXT, mangled);
	if (!f) { fprintf(stderr, "ref_tap: %s is not visible in the reference library\n", mangled); abort(); }
	return f;
}

} // namespace

extern "C" {
double ref_tap_seconds(int i) { return i >= 0 && i < S_COUNT ? g_secs[i] : 0.0; }
uint64_t ref_tap_calls(int i) { return i >= 0 && i < S_COUNT ? g_calls[i] : 0; }
void ref_tap_reset(void) { std::lock_guard<std::mutex> lk(g_mu); for (auto &v : g_secs) v = 0; for (auto &v : g_calls) v = 0; }
/* geometry parity: [0] tile-components checked, [1] code blocks compared, [2] mismatches (block count or any rectangle / index) */
uint64_t ref_tap_geometry(int i) { return i == 0 ? g_geo_tilecomps : i == 1 ? g_geo_blocks : g_geo_mismatch; }
}

namespace grk {

#define TAP_MEMBER(idx, name, mangled) \
	bool TileProcessor::name() { \
		using Fn = bool (*)(TileProcessor*); \
		static Fn real = next<Fn>(mangled); \
		Timer t(idx); \
		return real(this); \
	}

TAP_MEMBER(S_DC_ENC, dc_level_shift_encode, "_ZN3grk13TileProcessor21dc_level_shift_encodeEv")
TAP_MEMBER(S_MCT_ENC, mct_encode, "_ZN3grk13TileProcessor10mct_encodeEv")
TAP_MEMBER(S_DWT_ENC, dwt_encode, "_ZN3grk13TileProcessor10dwt_encodeEv")
TAP_MEMBER(S_MCT_DEC, mct_decode, "_ZN3grk13TileProcessor10mct_decodeEv")
TAP_MEMBER(S_DC_DEC, dc_level_shift_decode, "_ZN3grk13TileProcessor21dc_level_shift_decodeEv")

/* t1_encode: first compare the product's block table with the tile structure the reference has just built */
bool TileProcessor::t1_encode() {
	using Fn = bool (*)(TileProcessor*);
	static Fn real = next<Fn>("_ZN3grk13TileProcessor9t1_encodeEv");
	for (uint32_t c = 0; c < tile->numcomps; ++c) {
		auto tilec = tile->comps + c;
		auto tccp = m_tcp->tccps + c;
		gb200_comp_params p;
		memset(&p, 0, sizeof(p));
		p.x0 = tilec->x0; p.y0 = tilec->y0; p.x1 = tilec->x1; p.y1 = tilec->y1;
		p.numres = tilec->numresolutions;
		p.cblkw_expn = tccp->cblkw; p.cblkh_expn = tccp->cblkh;
		for (uint32_t r = 0; r < p.numres; ++r) { p.prcw_expn[r] = tccp->prcw[r]; p.prch_expn[r] = tccp->prch[r]; }
		for (uint32_t b = 0; b < 3 * p.numres - 2; ++b) p.band_numbps[b] = 1;
		const uint64_t n = gb200_enumerate_blocks(&p, 0, nullptr, 0);
		std::vector<gb200_cblk_info> info(n ? n : 1);
		gb200_enumerate_blocks(&p, 0, info.data(), n);
		uint64_t i = 0, bad = 0;
		for (uint32_t resno = 0; resno < tilec->numresolutions; ++resno) {
			auto res = tilec->resolutions + resno;
			uint32_t gpw = 0, gph = 0;
			gb200_precinct_grid(&p, resno, &gpw, &gph);
			if (gpw != res->pw || gph != res->ph) bad++;
			for (uint32_t bandno = 0; bandno < res->numbands; ++bandno) {
				auto band = res->bands + bandno;
				for (uint32_t precno = 0; precno < res->pw * res->ph; ++precno) {
					auto prc = band->precincts + precno;
					for (uint32_t cblkno = 0; cblkno < prc->cw * prc->ch; ++cblkno, ++i) {
						auto cblk = prc->cblks.enc + cblkno;
						if (i >= n) { bad++; continue; }
						const gb200_cblk_info &b = info[i];
						if (b.x0 != cblk->x0 || b.y0 != cblk->y0 || b.x1 != cblk->x1 || b.y1 != cblk->y1 || b.resno != resno
								|| b.bandno != band->bandno || b.precno != precno || b.cblkno != cblkno)
							bad++;
					}
				}
			}
		}
		if (i != n) bad++;
		std::lock_guard<std::mutex> lk(g_mu);
		g_geo_tilecomps++;
		g_geo_blocks += i;
		g_geo_mismatch += bad;
	}
	Timer t(S_T1_ENC);
	return real(this);
}

bool TileProcessor::encode_tile(uint16_t tile_no, BufferedStream *p_stream, uint64_t *p_data_written, uint64_t max_length,
		grk_codestream_info *p_cstr_info) {
	using Fn = bool (*)(TileProcessor*, uint16_t, BufferedStream*, uint64_t*, uint64_t, grk_codestream_info*);
	static Fn real = next<Fn>("_ZN3grk13TileProcessor11encode_tileEtPNS_14BufferedStreamEPmmP20_grk_codestream_info");
	Timer t(S_ENCODE_TILE);
	return real(this, tile_no, p_stream, p_data_written, max_length, p_cstr_info);
}

/* the host-side rate allocation the GPU path leaves to the codec: the packet-length simulation every probe of the threshold
 * bisection runs (TileProcessor.cpp:371-506, t2/T2.cpp:131-193) -- how much of encode_tile it is */
bool T2::encode_packets_simulate(uint16_t tile_no, grk_tcd_tile *p_tile, uint32_t max_layers, uint64_t *p_data_written, uint64_t max_len,
		uint32_t tp_pos) {
	using Fn = bool (*)(T2*, uint16_t, grk_tcd_tile*, uint32_t, uint64_t*, uint64_t, uint32_t);
	static Fn real = next<Fn>("_ZN3grk2T223encode_packets_simulateEtPNS_12grk_tcd_tileEjPmmj");
	Timer t(S_SIMULATE);
	return real(this, tile_no, p_tile, max_layers, p_data_written, max_len, tp_pos);
}

bool TileProcessor::decode_tile(ChunkBuffer *src_buf, uint16_t tile_no) {
	using Fn = bool (*)(TileProcessor*, ChunkBuffer*, uint16_t);
	static Fn real = next<Fn>("_ZN3grk13TileProcessor11decode_tileEPNS_11ChunkBufferEt");
	Timer t(S_DECODE_TILE);
	return real(this, src_buf, tile_no);
}

bool Tier1::decodeCodeblocks(grk_tcp *tcp, uint16_t blockw, uint16_t blockh, std::vector<decodeBlockInfo*> *blocks) {
	using Fn = bool (*)(Tier1*, grk_tcp*, uint16_t, uint16_t, std::vector<decodeBlockInfo*>*);
	static Fn real = next<Fn>("_ZN3grk5Tier116decodeCodeblocksEPNS_7grk_tcpEttPSt6vectorIPNS_15decodeBlockInfoESaIS5_EE");
	Timer t(S_T1_DEC);
	return real(this, tcp, blockw, blockh, blocks);
}

bool Wavelet::decode(TileProcessor *p_tcd, TileComponent *tilec, uint32_t numres, uint8_t qmfbid) {
	using Fn = bool (*)(Wavelet*, TileProcessor*, TileComponent*, uint32_t, uint8_t);
	(void) sizeof(Fn);
	using SFn = bool (*)(TileProcessor*, TileComponent*, uint32_t, uint8_t);
	static SFn real = next<SFn>("_ZN3grk7Wavelet6decodeEPNS_13TileProcessorEPNS_13TileComponentEjh");
	Timer t(S_DWT_DEC);
	return real(p_tcd, tilec, numres, qmfbid);
}

} // namespace grk
