// K5: EBCOT Tier-1 decoder: two kernels over the same pass code, de-quantisation in a following pass.
//
//   T1Part1::decode / post_decode   T1Part1.cpp:135-329   segment concat, /2 or x stepsize, scatter
//   t1_decode_cblk                  t1.cpp:1038-1130      plane loop, pass order
//   sig / ref / cln pass            t1.cpp:381-441, 588-637, 784-870
//   MQ decoder                      mqc_dec.cpp:161-214, mqc_dec_inl.h:60-189
//
// Decoding a block is one serial chain: every decision selects the next context.  A warp whose 32 lanes
// co-operate on ONE block spends its issue slots on communication (round-1 kernel: 125 warp instructions per
// decision, 19 ms for configs[1]), so a chain is always run by one instruction stream, CPU style (the three
// passes with the four rows of a stripe column unrolled on constant bit positions, one flag word per column
// in a register), and the two kernels differ in how chains are laid onto warps:
//  * t1_decode_kernel<LANES>: a block belongs to one THREAD, LANES (1..8) blocks per warp.  A warp
//    instruction then advances up to LANES chains, as far as their scan positions agree: for batches (tens of
//    thousands of blocks) the machine is bound by warp instructions issued and eight chains per warp win.
//  * t1_decode_uniform_kernel: a block belongs to one WARP whose 32 lanes all execute the same chain on the
//    same values (see there).  For launches that fit the machine with a warp per block (one image).
// Common to both:
//  * The whole state of a block lives on chip.  One 32-bit word per stripe column holds the
//    significance of its 3x6 neighbourhood, the signs of its own column, the visited and refined
//    bits (the reference's flag word, t1.h:97-168, re-laid row major); a 64x64 block takes 4.2 KB
//    of shared memory, so ~48 blocks are resident per SM and configs[1] (6804 blocks) is a single
//    wave over 148 SMs.  The MQ probability state of the 19 contexts is kept as packed Table C.2
//    rows (Qe, both transitions, MPS) so a decision costs one shared load before the interval
//    arithmetic starts.
//  * A sample that turns significant updates its neighbours' words in shared memory (result-less atomics /
//    plain read-modify-writes); magnitudes go straight to the coefficient plane: a store of the
//    mid-point value when the sample turns significant, a fire-and-forget RED.ADD of +-half a
//    step per refinement (t1.cpp:392-394, 485).  Nothing on the critical chain waits for memory
//    further away than shared.
//  * Compressed bytes are pulled through a 64-bit window with the next 8 bytes already in flight.
// t1_dec_clear_kernel zeroes the block areas first (the reference decodes into a zeroed tile
// buffer); t1_dec_finish_kernel applies /2 or x stepsize (T1Part1.cpp:300-327), one warp per block.
// The reference's artificial FF FF end marker (mqc_dec.cpp:161-177) is emulated by reading 0xFF
// past the end of the segment.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include "t1_tables.cuh"

namespace gb {

#ifndef DT_LANES
#define DT_LANES 2          // code blocks (active lanes) per warp of a single-wave launch; 1, 2, 4 or 8
#endif
// a launch with few blocks per SM (one small image) is bound by the latency of one block's chain: then every block gets
// a warp of its own, so that no two chains share an instruction stream
#ifndef DT_SPARSE_BLOCKS_PER_SM
#define DT_SPARSE_BLOCKS_PER_SM 16
#endif
// 1: launches that fit the machine with one warp per block go to the warp-uniform kernel
#ifndef DT_UNIFORM
#define DT_UNIFORM 1
#endif
// the warp-uniform kernel: CTAs per SM it is compiled for and threads per CTA (2 x 768: 40 registers, 48 blocks per SM)
#ifndef DU_MINB
#define DU_MINB 2
#endif
#ifndef DU_MAX_THREADS
#define DU_MAX_THREADS 768
#endif
// CTAs per SM the decode kernel is compiled for (register cap) and threads per CTA: 1 x 1024 (64 registers) or 2 x 768 (40)
#ifndef DT_MINB
#define DT_MINB 1
#endif
constexpr int DT_MAX_THREADS = DT_MINB == 1 ? 1024 : 768;
constexpr int DT_CTX_WORDS = 20;  // 19 context rows per block, padded
constexpr int DT_FIXED_WORDS = 96 + 512 + 64; // MQ table, zero-coding table (4 x 512 B), sign table (256 B)
#ifndef DU_MR_LIST
#define DU_MR_LIST 1
#endif
constexpr int DU_RACC_WORDS = 16 + 32; // warp-uniform decoder, per column of a stripe: one byte of refinement results, one 16-bit refinement code

// stripe-column word: bit 3r+j = significance of row r-1 (r = 0..5), column j-1 (j = 0 west, 1 own, 2 east);
// bit 18+r = sign of own column row r-1; bit 24+k = visited (k = 0..3); bit 28+k = refined before
__device__ __forceinline__ constexpr uint32_t fsig(int r, int j) { return 1u << (3 * r + j); }
constexpr uint32_t F_PI_ALL = 0xFu << 24;
constexpr uint32_t F_SIGMA_ALL = 0x3FFFFu;

// bits 4,7,10,13 (significance of the own column, rows 0..3) gathered into a nibble
__device__ __forceinline__ uint32_t own_sig4(uint32_t f) {
	return ((f >> 4) & 1u) | ((f >> 6) & 2u) | ((f >> 8) & 4u) | ((f >> 10) & 8u);
}
// rows whose 8-neighbourhood holds a significant sample
__device__ __forceinline__ uint32_t nbr4(uint32_t f) {
	return ((f & 0x1EFu) ? 1u : 0u) | ((f & (0x1EFu << 3)) ? 2u : 0u) | ((f & (0x1EFu << 6)) ? 4u : 0u) | ((f & (0x1EFu << 9)) ? 8u : 0u);
}

// ---- MQ decoder of one thread ------------------------------------------------------------------------
struct MqT {
	uint32_t a, c;   // A is kept in the high half-word (A << 16), like the Chigh half of C it is compared with
	int ct;
	uint32_t cur;    // byte at the read position
	uint64_t win;    // the bytes that follow it, next one lowest
	int nwin;        // bytes left in win (never 0 between calls)
	uint64_t pre;    // the 8 bytes after the window, already loaded
	const unsigned long long *base8; // 8-byte aligned address at or below the first byte
	uint32_t wi;     // index of the next aligned word to load
	uint32_t endoff; // offset of the end of the segment relative to base8
};

// aligned word j of the segment; bytes at or beyond its end read as 0xFF.  Called once per 8 consumed bytes (every ~80
// decisions) from 23 inlined decoders: kept out of line, with plain value arguments, so that it does not multiply the
// kernel's code size (instruction fetch is a visible stall of this kernel)
#ifdef DT_INLINE_REFILL
__device__ __forceinline__
#else
__device__ __noinline__
#endif
uint64_t mq_word_at(const unsigned long long *base8, uint32_t j, uint32_t endoff) {
	const uint32_t lo = 8u * j;
	if (lo >= endoff) return ~0ull;
	uint64_t w = __ldg(base8 + j);
	const uint32_t valid = endoff - lo;
	if (valid < 8u) w |= ~0ull << (8u * valid);
	return w;
}
__device__ __forceinline__ uint64_t mq_word(const MqT &q, uint32_t j) { return mq_word_at(q.base8, j, q.endoff); }

__device__ __forceinline__ void mq_advance(MqT &q) {
	q.win >>= 8;
	if (--q.nwin == 0) {
		q.win = q.pre;
		q.nwin = 8;
		q.pre = mq_word(q, q.wi++);
	}
}

// BYTEIN, mqc_dec_inl.h:114-134
__device__ __forceinline__ void mq_bytein(MqT &q) {
	const uint32_t next = (uint32_t) q.win & 0xFFu;
	if (q.cur == 0xFFu && next > 0x8Fu) { q.c += 0xFF00u; q.ct = 8; }
	else {
		const uint32_t ff = q.cur == 0xFFu ? 1u : 0u;
		q.c += next << (8 + ff);
		q.ct = 8 - (int) ff;
		q.cur = next;
		mq_advance(q);
	}
}

// INITDEC, mqc_dec.cpp:179-201
__device__ __forceinline__ void mq_init(MqT &q, const uint8_t *buf, uint32_t len) {
	const uint32_t off0 = (uint32_t) (reinterpret_cast<uintptr_t>(buf) & 7u);
	q.base8 = reinterpret_cast<const unsigned long long*>(buf - off0);
	q.endoff = off0 + len;
	q.win = mq_word(q, 0) >> (8u * off0);
	q.nwin = 8 - (int) off0;
	q.pre = mq_word(q, 1);
	q.wi = 2;
	q.cur = (uint32_t) q.win & 0xFFu;
	mq_advance(q);
	q.c = q.cur << 16;
	mq_bytein(q);
	q.c <<= 7;
	q.ct -= 7;
	q.a = 0x80000000u;
}

// context rows: qe << 16 | next(LPS) << 9 | next(MPS) << 2 | mps; the successors are byte offsets into the 94-entry
// table of (state, mps) pairs (so the SWITCH column of Table C.2 is folded into them), every field is one AND away.
// DECODE + RENORMD, mqc_dec_inl.h:60-86, 136-169
__device__ __forceinline__ uint32_t mq_decode(MqT &q, uint32_t *crow, const uint32_t *tab) {
	const uint32_t row = *crow;
	const uint32_t qs = row & 0xFFFF0000u, mps = row & 1u;
	const uint32_t a = q.a - qs, cs = q.c - qs;
	const bool lpsint = q.c < qs; // (C >> 16) < Qe : the LPS sub-interval
	if (!lpsint && (a & 0x80000000u)) { // MPS, no renormalisation: the one early exit
		q.a = a;
		q.c = cs;
		return mps;
	}
	// the remaining three cases with selects instead of branches: LPS sub-interval (conditional exchange when A < Qe),
	// or the MPS sub-interval with A below 0x8000 (exchange when A < Qe)
	const bool lps = (a < qs) != lpsint;
	q.c = lpsint ? q.c : cs;
	q.a = lpsint ? qs : a;
	*crow = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(tab) + (lps ? (row >> 7) & 0x1FCu : row & 0x1FCu));
	int sh = __clz(q.a);
	q.a <<= sh;
	if (sh > q.ct) {
		do { // RENORMD with BYTEIN whenever the bit counter runs out
			q.c <<= q.ct;
			sh -= q.ct;
			mq_bytein(q);
		} while (sh > q.ct);
	}
	q.c <<= sh;
	q.ct -= sh;
	return mps ^ (lps ? 1u : 0u);
}

// ---- the three coding passes of one stripe column, rows unrolled with constant bit positions ---------
struct Blk {
	MqT q;
	uint32_t *C;          // the 19 context rows of this block
	const uint32_t *tab;  // 94 (state, mps) rows
	const uint8_t *zc;    // zero-coding context by the 9 neighbourhood bits, this block's orientation
	const uint8_t *sc;    // sign context | xor bit << 5, by (N W E S significance at bits 1 3 5 7, signs at bits 0 2 4 6)
	int32_t *dst;
	uint32_t stride;
	int fw, nstripes;
	// code-block style switches only (STY kernels)
	bool vsc;             // stripe-causal contexts: a sample of row 0 is not shown to the stripe above (t1.cpp:177-182)
	const uint8_t *rbuf;  // raw (bypass) segment reader, mqc_dec_inl.h:90-112
	uint32_t rpos, rlen, rc;
	int rct;
	uint32_t lane;        // warp-uniform decoder only
	uint8_t *racc;        // warp-uniform decoder only: per column of the stripe, the refinements decoded (ref_row)
};

// one raw bit; bytes past the segment read as 0xFF, the byte after 0xFF carries 7 bits
__device__ __forceinline__ uint32_t raw_bit(Blk &b) {
	if (b.rct == 0) {
		const uint32_t nb = b.rpos < b.rlen ? b.rbuf[b.rpos] : 0xFFu;
		if (b.rc == 0xFFu) {
			if (nb > 0x8Fu) { b.rc = 0xFFu; b.rct = 8; }
			else { b.rc = nb; b.rpos++; b.rct = 7; }
		} else { b.rc = nb; b.rpos++; b.rct = 8; }
	}
	b.rct--;
	return (b.rc >> b.rct) & 1u;
}

constexpr uint32_t F_OWNSIG = 0x2490u; // significance of the own column, rows 0..3

// UNI: the warp-uniform decoder (t1_decode_uniform_kernel): all 32 lanes of a warp walk the SAME block with the same
// values, so a neighbour update is a plain read-modify-write (every lane writes the same word) and only lane 0 adds to
// the coefficient plane.  Otherwise (one thread per block): result-less shared atomics, one instruction each.
template<bool UNI>
__device__ __forceinline__ void flag_or(uint32_t *p, uint32_t v) {
	if (UNI) *p |= v;
	else atomicOr(p, v);
}

// sign context of row K (t1.cpp:115-140): context | xor bit << 5
template<int K>
__device__ __forceinline__ uint32_t sign_ctx(const Blk &b, uint32_t f, const uint32_t *cw) {
	const uint32_t fW = cw[-1], fE = cw[1];
	const uint32_t idx = ((f >> (3 * K)) & 0xAAu) | ((f >> (18 + K)) & 1u) | ((fW >> (17 + K)) & 4u) | ((fE >> (15 + K)) & 0x10u)
			| ((f >> (14 + K)) & 0x40u);
	return b.sc[idx];
}

template<int K, bool STY, bool UNI> __device__ __forceinline__ void mark(Blk &b, uint32_t &f, uint32_t *cw, int s, uint32_t off, int32_t oph, uint32_t neg);

// sign of a sample that just turned significant, mid-point store, neighbour updates (t1.cpp:168-195)
template<int K, bool STY = false, bool RAW = false, bool UNI = false>
__device__ __forceinline__ void sign_and_mark(Blk &b, uint32_t &f, uint32_t *cw, int s, uint32_t off, int32_t oph) {
	uint32_t neg;
	if (RAW) neg = raw_bit(b); // raw passes carry the sign itself (t1.cpp:233-260)
	else {
		const uint32_t v = sign_ctx<K>(b, f, cw);
		neg = mq_decode(b.q, b.C + (v & 31u), b.tab) ^ (v >> 5);
	}
	mark<K, STY, UNI>(b, f, cw, s, off, oph, neg);
}


template<int K, bool STY, bool UNI>
__device__ __forceinline__ void mark(Blk &b, uint32_t &f, uint32_t *cw, int s, uint32_t off, int32_t oph, uint32_t neg) {
	f |= fsig(K + 1, 1) | (neg << (19 + K));
	b.dst[off + K * b.stride] = neg ? -oph : oph;
	// the sample is the east neighbour of column x-1 and the west neighbour of column x+1
	flag_or<UNI>(cw - 1, fsig(K + 1, 2));
	flag_or<UNI>(cw + 1, fsig(K + 1, 0));
	if (K == 0 && s > 0 && !(STY && b.vsc)) { // row 4 of the stripe above
		uint32_t *up = cw - b.fw;
		flag_or<UNI>(up - 1, fsig(5, 2));
		flag_or<UNI>(up, fsig(5, 1) | (neg << 23));
		flag_or<UNI>(up + 1, fsig(5, 0));
	}
	if (K == 3 && s + 1 < b.nstripes) { // row -1 of the stripe below
		uint32_t *dn = cw + b.fw;
		flag_or<UNI>(dn - 1, fsig(0, 2));
		flag_or<UNI>(dn, fsig(0, 1) | (neg << 18));
		flag_or<UNI>(dn + 1, fsig(0, 0));
	}
}

// significance propagation, t1.cpp:381-441
template<int K, bool STY = false, bool RAW = false, bool UNI = false>
__device__ __forceinline__ void sig_row(Blk &b, uint32_t &f, uint32_t *cw, int s, uint32_t off, int32_t oph) {
	if ((f & (fsig(K + 1, 1) | (1u << (24 + K)))) == 0 && (f & (0x1EFu << (3 * K))) != 0) {
		const uint32_t d = RAW ? raw_bit(b) : mq_decode(b.q, b.C + b.zc[(f >> (3 * K)) & 0x1FFu], b.tab);
		f |= 1u << (24 + K);
		if (d) sign_and_mark<K, STY, RAW, UNI>(b, f, cw, s, off, oph);
	}
}

// magnitude refinement: +-half a step towards the decoded bit, t1.cpp:476-496, 588-637
// UNI: the four rows of a column only collect (member, decoded bit ^ sign) in `acc`; ref_flush then lets lane k add to row k
template<int K, bool RAW = false, bool UNI = false>
__device__ __forceinline__ void ref_row(Blk &b, uint32_t &f, uint32_t off, int32_t half, uint32_t &acc) {
	if ((f & (fsig(K + 1, 1) | (1u << (24 + K)))) == fsig(K + 1, 1)) {
		const uint32_t cx = (f & (1u << (28 + K))) ? CTX_MR0 + 2 : (f & (0x1EFu << (3 * K))) ? CTX_MR0 + 1 : CTX_MR0;
		const uint32_t d = RAW ? raw_bit(b) : mq_decode(b.q, b.C + cx, b.tab);
		if (UNI) acc |= (0x10u | (d ^ (f >> (19 + K)) & 1u)) << K; // bit 4 + K: member, bit K: up (1) or down (0)
		else {
			const uint32_t neg = (f >> (19 + K)) & 1u;
			atomicAdd(b.dst + off + K * b.stride, (d ^ neg) ? half : -half);
		}
		f |= 1u << (28 + K);
	}
}
// end of a stripe of a refinement pass: lane l looks at columns l and l + 32 and adds to the rows that were refined
// (fire-and-forget REDs, the one divergent region of the pass)
// Out of line on purpose: inlined (and unrolled) into the pass code it cost 16 % of the kernel's time (configs[1]: 9.5 ms
// against 8.2).
__device__ __noinline__ void ref_flush(uint8_t *racc, int32_t *dst, uint32_t stride, uint32_t lane, int w, uint32_t off0, int32_t half) {
	#pragma unroll 1
	for (int x = (int) lane; x < w; x += 32) {
		const uint32_t acc = racc[x];
		if (!acc) continue;
		racc[x] = 0;
		int32_t *p = dst + off0 + x;
		#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (acc & (0x10u << k)) atomicAdd(p + k * stride, (acc >> k & 1u) ? half : -half);
	}
}

// cleanup, t1.cpp:784-870; start / implied: the row whose 1 the run-length code already delivered
template<int K, bool STY = false, bool UNI = false>
__device__ __forceinline__ void cln_row(Blk &b, uint32_t &f, uint32_t *cw, int s, uint32_t off, int32_t oph, int start, bool implied) {
	if (K >= start && (f & (fsig(K + 1, 1) | (1u << (24 + K)))) == 0) {
		uint32_t d = 1;
		if (!(implied && K == start)) d = mq_decode(b.q, b.C + b.zc[(f >> (3 * K)) & 0x1FFu], b.tab);
		if (d) sign_and_mark<K, STY, false, UNI>(b, f, cw, s, off, oph);
	}
}

// one coding pass over the whole block
template<bool STY, bool RAW, bool UNI = false>
__device__ __forceinline__ void run_pass(Blk &b, uint32_t *F, int w, int fw, int type, int bp1, uint32_t last_pi) {
	// values carry one extra low bit: the plane weight is 1 << bp1, the mid-point sits half a step above
	const int32_t half = (1 << bp1) >> 1, oph = (1 << bp1) | half;
	for (int s = 0; s < b.nstripes; ++s) {
		uint32_t *cw = F + s * fw + 1;
		uint32_t off = (uint32_t) (4 * s) * b.stride;
		if (type == 0) {
			for (int x = 0; x < w; ++x, ++cw, ++off) {
				uint32_t f = *cw;
				if (!(f & F_SIGMA_ALL) || (f & F_OWNSIG) == F_OWNSIG) continue; // nothing significant around / nothing left to find
				sig_row<0, STY, RAW, UNI>(b, f, cw, s, off, oph);
				sig_row<1, STY, RAW, UNI>(b, f, cw, s, off, oph);
				sig_row<2, STY, RAW, UNI>(b, f, cw, s, off, oph);
				sig_row<3, STY, RAW, UNI>(b, f, cw, s, off, oph);
				*cw = f;
			}
		} else if (type == 1 && UNI && DU_MR_LIST) {
			// Warp-uniform decoder: which samples a refinement pass codes, and in which of the three contexts, is fixed when
			// the pass starts (nothing turns significant in it).  The 32 lanes look at 32 columns of the stripe at once: members
			// (significant, not visited), their contexts and signs go into one 16-bit code per column, the refined bits are set
			// on the spot, and a ballot says which columns have members at all; the serial walk then touches only those
			// columns and does nothing per sample but decode.
			uint16_t *const codes = reinterpret_cast<uint16_t*>(b.racc + 64);
			for (int x0 = 0; x0 < w; x0 += 32) {
				const int xl = x0 + (int) b.lane;
				uint32_t code = 0;
				if (xl < w) {
					const uint32_t f = cw[xl];
					const uint32_t m4 = own_sig4(f) & ~(f >> 24) & 0xFu;
					if (m4) {
						const uint32_t refd = f >> 28, nb = nbr4(f);
						uint32_t ctx = 0; // two bits per row: 2 = refined before, 1 = a significant neighbour, 0 = neither (t1.cpp:476-496)
						#pragma unroll
						for (int k = 0; k < 4; ++k) ctx |= ((refd >> k & 1u) ? 2u : (nb >> k & 1u)) << (2 * k);
						code = m4 | ctx << 4 | ((f >> 19) & 0xFu) << 12;
						cw[xl] = f | m4 << 28;
					}
					codes[xl] = (uint16_t) code;
				}
				uint32_t todo = __ballot_sync(0xffffffffu, code != 0);
				while (todo) {
					const int x = x0 + __ffs(todo) - 1;
					todo &= todo - 1;
					const uint32_t cd = codes[x];
					uint32_t acc = (cd & 0xFu) << 4;
					#pragma unroll
					for (int k = 0; k < 4; ++k)
						if (cd >> k & 1u) {
							const uint32_t d = RAW ? raw_bit(b) : mq_decode(b.q, b.C + CTX_MR0 + ((cd >> (4 + 2 * k)) & 3u), b.tab);
							acc |= ((d ^ (cd >> (12 + k))) & 1u) << k;
						}
					b.racc[x] = (uint8_t) acc;
				}
			}
			ref_flush(b.racc, b.dst, b.stride, b.lane, w, (uint32_t) (4 * s) * b.stride, half);
		} else if (type == 1) {
			for (int x = 0; x < w; ++x, ++cw, ++off) {
				uint32_t f = *cw;
				if (!(f & F_OWNSIG)) continue;
				uint32_t acc = 0;
				ref_row<0, RAW, UNI>(b, f, off, half, acc);
				ref_row<1, RAW, UNI>(b, f, off, half, acc);
				ref_row<2, RAW, UNI>(b, f, off, half, acc);
				ref_row<3, RAW, UNI>(b, f, off, half, acc);
				if (UNI && acc) b.racc[x] = (uint8_t) acc;
				*cw = f;
			}
			if (UNI) ref_flush(b.racc, b.dst, b.stride, b.lane, w, (uint32_t) (4 * s) * b.stride, half);
		} else if (!RAW) {
			const uint32_t keep_pi = s == b.nstripes - 1 ? last_pi : 0u;
			for (int x = 0; x < w; ++x, ++cw, ++off) {
				uint32_t f = *cw;
				int start = 0;
				bool implied = false;
				if ((f & (F_PI_ALL | F_SIGMA_ALL)) == 0) { // run-length mode, t1.cpp:749 (full stripes only: keep_pi)
					if (!mq_decode(b.q, b.C + CTX_AGG, b.tab)) continue;
					start = (int) mq_decode(b.q, b.C + CTX_UNI, b.tab) << 1;
					start |= (int) mq_decode(b.q, b.C + CTX_UNI, b.tab);
					implied = true;
				}
				cln_row<0, STY, UNI>(b, f, cw, s, off, oph, start, implied);
				cln_row<1, STY, UNI>(b, f, cw, s, off, oph, start, implied);
				cln_row<2, STY, UNI>(b, f, cw, s, off, oph, start, implied);
				cln_row<3, STY, UNI>(b, f, cw, s, off, oph, start, implied);
				*cw = (f & ~F_PI_ALL) | keep_pi;
			}
		}
	}
}

// the coding passes of one block in order: cln(numbps), then sig / ref / cln per lower plane (t1.cpp:1067-1114)
template<bool STY, bool UNI>
__device__ __forceinline__ void run_block(Blk &b, uint32_t *F, const uint32_t *tab, const DecBlock &B, const DecInput &I, uint32_t bid,
		const uint8_t *__restrict__ data, int w, int fw, int top, int numbps, uint32_t last_pi,
		const uint32_t *__restrict__ seg_start, const DecSeg *__restrict__ segs) {
	int bp1 = top, type = 2;
	if (!STY) {
		mq_init(b.q, data + I.data_offset, I.data_len);
		const int npass = min((int) I.numpasses, 3 * top - 2);
		for (int pass = 0; pass < npass; ++pass) {
			run_pass<false, false, UNI>(b, F, w, fw, type, bp1, last_pi);
			if (++type == 3) { type = 0; bp1--; }
		}
	} else {
		// codeword segments as Tier-2 delivers them (t1.cpp:1067-1114): every segment restarts the MQ decoder, or the raw
		// reader for the bypassed passes of LAZY; RESET clears the contexts after every MQ pass; SEGSYM appends four
		// UNIFORM decisions to every cleanup pass
		const uint32_t sty = B.sty;
		b.vsc = (sty & STY_VSC) != 0;
		const uint32_t s0 = seg_start ? seg_start[bid] : 0u, nsegs = seg_start ? seg_start[bid + 1] - s0 : 1u;
		uint32_t off = 0;
		for (uint32_t sg = 0; sg < nsegs && bp1 >= 1; ++sg) {
			// a segment never reaches past the block's bytes (corrupt packet headers: sum of the lengths > data_len); what is
			// missing reads as the 0xFF fill of an exhausted segment
			const uint32_t left = off < I.data_len ? I.data_len - off : 0u;
			const uint32_t len = min(seg_start ? segs[s0 + sg].len : I.data_len, left), np = seg_start ? segs[s0 + sg].numpasses : I.numpasses;
			const bool raw = (sty & STY_LAZY) && bp1 <= numbps - 4 && type < 2;
			const uint8_t *seg = data + I.data_offset + off;
			if (raw) { b.rbuf = seg; b.rpos = 0; b.rlen = len; b.rc = 0; b.rct = 0; } // mqc_raw_init_dec
			else mq_init(b.q, seg, len);
			off += len;
			for (uint32_t p = 0; p < np && bp1 >= 1; ++p) {
				if (raw) run_pass<true, true, UNI>(b, F, w, fw, type, bp1, last_pi);
				else {
					run_pass<true, false, UNI>(b, F, w, fw, type, bp1, last_pi);
					if (type == 2 && (sty & STY_SEGSYM)) // t1_dec_clnpass_check_segsym: read, a mismatch only warns
						for (int i = 0; i < 4; ++i) mq_decode(b.q, b.C + CTX_UNI, b.tab);
					if (sty & STY_RESET) {
						#pragma unroll
						for (int i = 0; i < NCTX; ++i) b.C[i] = tab[2 * (i == CTX_ZC0 ? 4 : i == CTX_AGG ? 3 : i == CTX_UNI ? 46 : 0)];
					}
				}
				if (++type == 3) { type = 0; bp1--; }
			}
		}
	}
}

template<int LANES, bool STY>
__global__ void __launch_bounds__(DT_MAX_THREADS, DT_MINB) t1_decode_kernel(const DecBlock *__restrict__ blocks,
		const DecInput *__restrict__ inputs, uint32_t nblocks, const uint8_t *__restrict__ data, int fw, int fwords, int nslots,
		const uint32_t *__restrict__ seg_start, const DecSeg *__restrict__ segs) {
	extern __shared__ __align__(16) uint32_t sm[];
	uint32_t *tab = sm;                                        // 94 (state, mps) rows
	uint8_t *Lzc = reinterpret_cast<uint8_t*>(sm + 96);        // zero-coding context by the 9 neighbourhood bits of a word
	uint8_t *Lsc = Lzc + 2048;
	uint32_t *flags = sm + DT_FIXED_WORDS;                     // [slot][fwords]
	uint32_t *ctxrows = flags + (size_t) nslots * fwords;      // [slot][DT_CTX_WORDS]
	for (int i = threadIdx.x; i < 94; i += blockDim.x) {
		const uint32_t r = c_mq[i >> 1], mps = i & 1u, sw = (r >> 28) & 1u;
		const uint32_t nm = ((r >> 16) & 63u) * 2u + mps, nl = ((r >> 22) & 63u) * 2u + (mps ^ sw);
		tab[i] = (r << 16) | (nl << 9) | (nm << 2) | mps;
	}
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
		const int o = i >> 9, n9 = i & 511;
		const int idx8 = (n9 & 7) | ((n9 >> 3) & 1) << 3 | ((n9 >> 5) & 1) << 4 | ((n9 >> 6) & 7) << 5;
		Lzc[i] = c_zc[o][idx8];
	}
	for (int i = threadIdx.x; i < 256; i += blockDim.x) {
		const int orig = (i >> 1 & 1) | (i >> 3 & 1) << 1 | (i >> 5 & 1) << 2 | (i >> 7 & 1) << 3
				| (i & 1) << 4 | (i >> 2 & 1) << 5 | (i >> 4 & 1) << 6 | (i >> 6 & 1) << 7;
		Lsc[i] = c_sc[orig];
	}
	__syncthreads();

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane >= LANES) return;
	const int slot = warp * LANES + lane;
	const uint32_t bid = blockIdx.x * (uint32_t) nslots + (uint32_t) slot;
	if (slot >= nslots || bid >= nblocks) return;
	const DecBlock B = blocks[bid];
	const DecInput I = inputs[bid];
	const int w = B.w, h = B.h;
	const int numbps = (int) I.numbps;
	// T1Part1.cpp:139, t1.cpp:1055-1060: nothing to decode; the cleared block area stays zero.  With an ROI up-shift the
	// first coded plane lies roishift planes higher (bpno_plus_one = roishift + numbps)
	const int top = numbps + (int) B.roishift;
	if (I.numpasses == 0 || I.data_len == 0 || top == 0 || top > 30 || w == 0 || h == 0) return;

	Blk b;
	b.nstripes = (h + 3) >> 2;
	b.fw = fw;
	uint32_t *F = flags + (size_t) slot * fwords; // word of stripe s, column x: F[s * fw + x + 1]
	for (int i = 0; i < b.nstripes * fw; ++i) F[i] = 0;
	// rows below the block in a partial last stripe count as visited for ever (t1.cpp:990-1003)
	const uint32_t last_pi = (0xFu << (h - 4 * (b.nstripes - 1)) & 0xFu) << 24;
	if (last_pi)
		for (int x = 0; x < w; ++x) F[(b.nstripes - 1) * fw + 1 + x] = last_pi;
	b.C = ctxrows + slot * DT_CTX_WORDS;
	b.tab = tab;
	#pragma unroll
	for (int i = 0; i < NCTX; ++i) b.C[i] = tab[2 * (i == CTX_ZC0 ? 4 : i == CTX_AGG ? 3 : i == CTX_UNI ? 46 : 0)]; // mqc_dec.cpp:207-214
	b.zc = Lzc + 512 * B.orient;
	b.sc = Lsc;
	b.dst = B.dst;
	b.stride = B.stride;
	run_block<STY, false>(b, F, tab, B, I, bid, data, w, fw, top, numbps, last_pi, seg_start, segs);
}

// ---- warp-uniform decoder: one WARP per code block, every lane running the same chain --------------------
// For launches that cannot fill the machine with thread-per-block chains (one image: 6 804 blocks, 11.5 per scheduler).
// All 32 lanes of a warp decode the SAME block redundantly: every address and every branch condition derives from
// blockIdx, a warp index taken through a shuffle, and loads at warp-uniform addresses, so the compiler sees uniform control
// flow and emits no reconvergence bookkeeping (BSSY / BSYNC / BREAK: 10 of the 65 instructions per decision of the kernel
// above), and twice as many warps per scheduler hide the latency of a chain (issue slots busy 79 % against 66 %).  An
// instruction costs an issue slot and a pass through the pipe whether one lane or 32 are live, so the redundant lanes are
// free.  The lanes differ in the refinement pass only.  Before a stripe is walked they look at 32 columns at a time and write,
// per column, which rows the pass codes, in which context and with which sign (membership is fixed when the pass starts), so
// the serial walk visits only columns with members and does nothing per sample but decode; it notes per column which rows it
// refined and in which direction (one shared-memory byte), and at the end of the stripe ref_flush lets the lanes add the half
// steps to the coefficient plane for 32 columns at a time -- per refinement that replaces the row tests, the context
// selection, address arithmetic, a predicated RED and its reconvergence region on the serial chain by a handful of
// instructions.  Two CTAs of up to 24 warps per SM (40 registers).
template<bool STY>
__global__ void __launch_bounds__(DU_MAX_THREADS, DU_MINB) t1_decode_uniform_kernel(const DecBlock *__restrict__ blocks,
		const DecInput *__restrict__ inputs, uint32_t nblocks, const uint8_t *__restrict__ data, int fw, int fwords,
		const uint32_t *__restrict__ seg_start, const DecSeg *__restrict__ segs) {
	extern __shared__ __align__(16) uint32_t sm[];
	uint32_t *tab = sm;                                        // 94 (state, mps) rows
	uint8_t *Lzc = reinterpret_cast<uint8_t*>(sm + 96);        // zero-coding context by the 9 neighbourhood bits of a word
	uint8_t *Lsc = Lzc + 2048;
	for (int i = threadIdx.x; i < 94; i += blockDim.x) {
		const uint32_t r = c_mq[i >> 1], mps = i & 1u, sw = (r >> 28) & 1u;
		const uint32_t nm = ((r >> 16) & 63u) * 2u + mps, nl = ((r >> 22) & 63u) * 2u + (mps ^ sw);
		tab[i] = (r << 16) | (nl << 9) | (nm << 2) | mps;
	}
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
		const int o = i >> 9, n9 = i & 511;
		const int idx8 = (n9 & 7) | ((n9 >> 3) & 1) << 3 | ((n9 >> 5) & 1) << 4 | ((n9 >> 6) & 7) << 5;
		Lzc[i] = c_zc[o][idx8];
	}
	for (int i = threadIdx.x; i < 256; i += blockDim.x) {
		const int orig = (i >> 1 & 1) | (i >> 3 & 1) << 1 | (i >> 5 & 1) << 2 | (i >> 7 & 1) << 3
				| (i & 1) << 4 | (i >> 2 & 1) << 5 | (i >> 4 & 1) << 6 | (i >> 6 & 1) << 7;
		Lsc[i] = c_sc[orig];
	}
	__syncthreads();
	const int lane = threadIdx.x & 31;
	// the warp index as the compiler can see it is the same in every lane (a shuffle from lane 0): what follows is uniform
	const uint32_t wid = (uint32_t) __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0);
	// warp k of CTA c takes block c + k * gridDim.x: the blocks of one SM come from all over the image (the table is ordered
	// by tile, component, resolution), so every SM gets the same mix of heavy and light blocks
	const uint32_t bid = blockIdx.x + wid * gridDim.x;
	if (bid >= nblocks) return;
	uint32_t *F = sm + DT_FIXED_WORDS + wid * (uint32_t) (fwords + DT_CTX_WORDS + DU_RACC_WORDS); // word of stripe s, column x: F[s * fw + x + 1]
	const DecBlock B = blocks[bid];
	const DecInput I = inputs[bid];
	const int w = B.w, h = B.h;
	const int numbps = (int) I.numbps;
	const int top = numbps + (int) B.roishift;
	if (I.numpasses == 0 || I.data_len == 0 || top == 0 || top > 30 || w == 0 || h == 0) return;
	Blk b;
	b.nstripes = (h + 3) >> 2;
	b.fw = fw;
	b.lane = (uint32_t) lane;
	for (int i = lane; i < b.nstripes * fw; i += 32) F[i] = 0;
	__syncwarp();
	const uint32_t last_pi = (0xFu << (h - 4 * (b.nstripes - 1)) & 0xFu) << 24;
	if (last_pi)
		for (int x = lane; x < w; x += 32) F[(b.nstripes - 1) * fw + 1 + x] = last_pi;
	b.C = F + fwords;
	b.tab = tab;
	b.racc = reinterpret_cast<uint8_t*>(F + fwords + DT_CTX_WORDS);
	if (lane < 16) F[fwords + DT_CTX_WORDS + lane] = 0; // the result bytes; the codes are written before they are read
	#pragma unroll
	for (int i = 0; i < NCTX; ++i) b.C[i] = tab[2 * (i == CTX_ZC0 ? 4 : i == CTX_AGG ? 3 : i == CTX_UNI ? 46 : 0)]; // mqc_dec.cpp:207-214
	__syncwarp();
	b.zc = Lzc + 512 * B.orient;
	b.sc = Lsc;
	b.dst = B.dst;
	b.stride = B.stride;
	run_block<STY, true>(b, F, tab, B, I, bid, data, w, fw, top, numbps, last_pi, seg_start, segs);
}

// ---- before / after: one warp per block, coalesced ---------------------------------------------------
__global__ void __launch_bounds__(256) t1_dec_clear_kernel(const DecBlock *__restrict__ blocks, uint32_t nblocks) {
	const uint32_t bid = blockIdx.x * 8 + (threadIdx.x >> 5);
	if (bid >= nblocks) return;
	const int lane = threadIdx.x & 31;
	const DecBlock B = blocks[bid];
	for (int y = 0; y < B.h; ++y)
		for (int x = lane; x < B.w; x += 32) B.dst[(size_t) y * B.stride + x] = 0;
}

// T1Part1.cpp:230-252: samples at or above 2^roishift belong to the region of interest and are shifted back down;
// T1Part1.cpp:300-327: reversible /2 (C division), irreversible float(value) * stepsize
__global__ void __launch_bounds__(256) t1_dec_finish_kernel(const DecBlock *__restrict__ blocks, uint32_t nblocks) {
	const uint32_t bid = blockIdx.x * 8 + (threadIdx.x >> 5);
	if (bid >= nblocks) return;
	const int lane = threadIdx.x & 31;
	const DecBlock B = blocks[bid];
	for (int y = 0; y < B.h; ++y)
		for (int x = lane; x < B.w; x += 32) {
			int32_t *p = B.dst + (size_t) y * B.stride + x;
			int32_t v = *p;
			if (B.roishift) {
				const int32_t mag = abs(v);
				if (mag >= (1 << B.roishift)) v = v < 0 ? -(mag >> B.roishift) : (mag >> B.roishift);
			}
			*p = B.reversible ? v / 2 : __float_as_int(__fmul_rn((float) v, B.stepsize));
		}
}

int launch_t1_decode(const DecBlock *blocks, const DecInput *inputs, uint32_t nblocks, const uint8_t *data,
		uint32_t max_w, uint32_t max_h, int styles, const uint32_t *seg_start, const DecSeg *segs, cudaStream_t s) {
	if (!nblocks) return 0;
	ensure_t1_tables();
	int dev = 0, sms = 148, smem_max = 0, smem_sm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	const int smem_optin = smem_max;
	cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
	if (DT_MINB > 1) smem_max = std::min(smem_max, smem_sm / DT_MINB - 1024); // every resident CTA also reserves 1 KB
	if (max_w < 1) max_w = 1;
	if (max_h < 1) max_h = 1;
	const int fw = (int) max_w + 2;
	const int want = (int) ((nblocks + (uint32_t) (sms * DT_MINB) - 1) / (uint32_t) (sms * DT_MINB)); // spread a small job over every SM
	// flag words per block and block slots that fit the shared memory of a CTA, for `lanes` blocks per warp (the slots of one
	// warp start 32 / lanes banks apart)
	auto shape = [&](int lanes, int &fwords, int &cap) {
		fwords = (int) ((max_h + 3) / 4) * fw;
		fwords += fwords & 1;
		int bank0 = 32 / lanes % 32;
		if (bank0 & 1) bank0 = 2;
		while (fwords % 32 != bank0) fwords += 2;
		cap = (smem_max - DT_FIXED_WORDS * 4) / ((fwords + DT_CTX_WORDS) * 4);
	};
	// Blocks per warp.  A launch with few blocks per SM is bound by the latency of one block's chain: every block gets a warp of
	// its own.  A single wave (configs[1]: 46 blocks per SM) runs best with two blocks per warp.  A launch of many waves is bound by
	// warp instructions issued, and an instruction costs the same with one live lane or eight: as many blocks per warp as leave
	// about twenty warps resident per SM (measured, profiles/README.md: 30 cinema frames, 32x32 blocks, 189 slots per SM:
	// 68 ms at 2 per warp, 48 at 4, 41 at 8, 46 at 16; configs[2] planes, 64x64 blocks, 48 slots per SM: 133 ms at 2, 147 at 6).
	// Launches whose blocks all fit on the machine at once with a warp each (one image: configs[0] 7 blocks per SM, configs[1]
	// 46, one cinema frame 45) go to the warp-uniform kernel: DU_MINB CTAs per SM, each holding its half of the SM's share.
	// Measured (profiles/README.md): configs[1] 8.19 ms against 9.10 with two blocks per warp, configs[0] 4.92 against 6.07, one
	// cinema frame 4.00 against 4.57; batches (30 cinema frames: 84 ms against 41, configs[2]: 157 against 131) stay with the
	// thread-per-block kernel below, where eight chains share a warp's instruction stream.
	{
		const int ufwords = (int) ((max_h + 3) / 4) * fw;
		const int per_warp = (ufwords + DT_CTX_WORDS + DU_RACC_WORDS) * 4;
		const int ucap = std::min((smem_sm / DU_MINB - 1024 - DT_FIXED_WORDS * 4) / per_warp, DU_MAX_THREADS / 32); // warps (= blocks) per CTA
		int warps = (want + DU_MINB - 1) / DU_MINB;
		int uniform = DT_UNIFORM && warps <= ucap;
		if (const char *e = getenv("GB200_T1_DEC_UNIFORM")) uniform = atoi(e) != 0; // measurement knob
		if (uniform) {
			if (warps > ucap) warps = ucap; // forced by the knob: several waves
			if (warps < 1) return 1;
			const size_t smem = (size_t) DT_FIXED_WORDS * 4 + (size_t) warps * per_warp;
			const bool sty = styles || seg_start;
			auto kernel = sty ? t1_decode_uniform_kernel<true> : t1_decode_uniform_kernel<false>;
			// always the device's limit, never this launch's size: contexts on several host threads launch the same kernel with
			// different sizes, and a smaller value set by one thread would fail the other's launch
			if (smem > (size_t) smem_optin || cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) != cudaSuccess) return 1;
			t1_dec_clear_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(blocks, nblocks);
			kernel<<<(nblocks + warps - 1) / warps, warps * 32, smem, s>>>(blocks, inputs, nblocks, data, fw, ufwords, seg_start, segs);
			t1_dec_finish_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(blocks, nblocks);
			return 0;
		}
	}
	int lanes, fwords, cap;
	if (want * DT_MINB <= DT_SPARSE_BLOCKS_PER_SM) lanes = 1;
	else {
		lanes = DT_LANES;
		shape(lanes, fwords, cap);
		if (want > cap && !getenv("GB200_T1_DEC_LANES")) {
			for (int l : {8, 4}) {
				int f, c;
				shape(l, f, c);
				if (c >= 20 * l) { lanes = l; break; }
			}
		}
	}
	if (const char *e = getenv("GB200_T1_DEC_LANES")) { const int l = atoi(e); if (l == 1 || l == 2 || l == 4 || l == 8) lanes = l; } // measurement knob
	shape(lanes, fwords, cap);
	const int per_slot = (fwords + DT_CTX_WORDS) * 4;
	int nslots = cap < want ? cap : want;
	const int max_slots = DT_MAX_THREADS / 32 * lanes;
	if (nslots > max_slots) nslots = max_slots;
	nslots = (nslots + lanes - 1) / lanes * lanes;
	while (nslots > cap) nslots -= lanes;
	if (nslots < lanes) nslots = cap; // a partial warp: blocks too large for `lanes` of them
	if (nslots < 1) return 1;
	const size_t smem = (size_t) DT_FIXED_WORDS * 4 + (size_t) nslots * per_slot;
	const bool sty = styles || seg_start;
	auto kernel = lanes == 1 ? (sty ? t1_decode_kernel<1, true> : t1_decode_kernel<1, false>)
			: lanes == 2 ? (sty ? t1_decode_kernel<2, true> : t1_decode_kernel<2, false>)
			: lanes == 4 ? (sty ? t1_decode_kernel<4, true> : t1_decode_kernel<4, false>)
			: (sty ? t1_decode_kernel<8, true> : t1_decode_kernel<8, false>);
	if (smem > (size_t) smem_optin || cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin) != cudaSuccess) return 1; // see above
	const int threads = (nslots + lanes - 1) / lanes * 32;
	t1_dec_clear_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(blocks, nblocks);
	kernel<<<(nblocks + nslots - 1) / nslots, threads, smem, s>>>(blocks, inputs, nblocks, data, fw, fwords, nslots, seg_start, segs);
	t1_dec_finish_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(blocks, nblocks);
	return 0;
}

} // namespace gb
