// Warp-level affine-gap Needleman-Wunsch with traceback statistics: the device routine behind the alignment
// kernel (kmagpu_align.cu) and the stand-alone NW batch entry point.
//
// WHAT is computed is NW_score (nw.c:642-890) and NW_band_score (nw.c:892-1188): same int32 recurrences, same tie
// rules, same traceback byte per cell, same start-cell selection for the k modes, same walk. HOW is ours:
//
//   * Coordinates. The reference fills from the END of both sequences (m = t_len-1..0, n = q_len-1..0). We use
//     i = t_len-1-m, j = q_len-1-n, so cell (i,j) needs (i,j-1) [D,Q], (i-1,j) [D,P] and (i-1,j-1) [D].
//   * Band as geometry. The reference's skewed band buffer (row m covers query positions c_m-half..c_m+half, c_m
//     falling by one per row) is the column range [max(0,a+i), min(q_len-1,a+i+band)] of row i; the full matrix is
//     the same code with the range [0, q_len-1].
//   * Continuous wavefront. Lane l owns rows l, l+32, l+64, ... Row i = 32r+l computes its cell of in-row offset u
//     (u = j for the full matrix, u = j-i-a inside the band) at step T = s*l + P*r + u, s = 1 (full) or 2 (band),
//     P = max(W, 33s-...) the row period, W the row width. Every lane then needs, at every step, exactly the (D,P)
//     its upper neighbour lane produced one step earlier (one shuffle each) and the D it received the step before
//     (the diagonal) -- for the full matrix and for the band alike. Lane 0 is fed by lane 31 of the previous round
//     through a row buffer that it prefetches one step ahead. With W >= 33s the lanes never idle between rounds, so
//     a band of 65 columns runs at ~97 % lane utilisation instead of the ~50 % of a strip-by-strip wavefront.
//   * Traceback bytes are stored step-major / lane-minor: one coalesced 32-byte store per warp per step.
//
// The per-lane step and the finishing walk are plain functions of (geometry, lane state, neighbour values) so the
// same source is compiled by g++ into a lock-step emulator (tests/emu/nw_emu.cpp) that is checked against the
// oracle on the CPU; the CUDA wrapper at the bottom only adds the shuffles.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NW_HD __host__ __device__ __forceinline__
#else
#define NW_HD inline
#endif

#define NW_RING 256

struct NwPen { int W1, U, MM, M; int d[25]; int d8; };   // d8: every d[] fits a signed byte (per-row table in a register)
struct NwStat { int score, len, pos, match, tGaps, qGaps; };
struct alignas(8) NwRow { int D, P; };

struct NwGeo {
	int t_len, q_len, k, banded, band, a, W, P, s, R, Tmax, NEG, W1, U, rmask, d8;
	int C;   // row sweep: cells per lane (0: the continuous wavefront)
	NW_HD int off(int i) const { return banded ? a + i : 0; }
	NW_HD int jlo(int i) const { int v = banded ? a + i : 0; return v < 0 ? 0 : v; }
	NW_HD int jhi(int i) const { int v = banded ? a + i + band : q_len - 1; return v < q_len - 1 ? v : q_len - 1; }
	// D of the boundary column (all query consumed) beside row i, and of the boundary row (all template consumed)
	NW_HD int bcol(int i) const { return i < 0 ? 0 : (0 < k ? 0 : W1 + i * U); }
	NW_HD int brow(int j) const { return j < 0 ? 0 : (k == 2 ? 0 : W1 + j * U); }
	NW_HD size_t eaddr(int i, int j) const {
		if (C) return (size_t)i * (size_t)(32 * C) + (size_t)(j - off(i));   // row sweep: row-major, 32*C bytes per row
		const int l = i & 31;
		return ((size_t)(s * l + P * (i >> 5) + (j - off(i)))) * 32 + (size_t)l;
	}
	NW_HD size_t ebytes() const { return C ? (size_t)t_len * (size_t)(32 * C) + 8 : (size_t)Tmax * 32; }
};

// false: geometry the reference itself never produces (band narrower than the length difference)
#ifndef NW_RS_MAXC
#define NW_RS_MAXC 8
#endif
#ifndef NW_RS_FULL
#define NW_RS_FULL 1   // row sweep for full matrices
#endif
#ifndef NW_RS_BAND
#define NW_RS_BAND 1   // row sweep for bands
#endif
// rs: the widest row (in cells per lane, at most NW_RS_MAXC) the caller's kernel sweeps row by row; 0: none. Inside the alignment kernels every
// variant of the row sweep measured slower than the wavefront (profiles/r01_ab_rowsweep.log: those kernels stall on
// instruction fetch, and the sweep adds a dozen more loops to them), so they pass false.
NW_HD bool nw_geo_init(NwGeo &g, const NwPen &pen, int t_len, int q_len, int k, int band, int rs) {
	g.t_len = t_len; g.q_len = q_len; g.k = k; g.banded = band != 0;
	g.W1 = pen.W1; g.U = pen.U;
	if (band & 1) ++band;   // nw.c:374-376
	g.band = band;
	if (g.banded) {
		const int half = band >> 1, c0 = (t_len + q_len) >> 1;
		int dlt = t_len - q_len;
		if (dlt < 0) dlt = -dlt;
		if (band < dlt + 2 || band < 2) return false;
		g.a = q_len - 1 - c0 - half;
		g.W = band + 1; g.s = 2; g.P = g.W < 65 ? 65 : g.W;
	} else {
		g.a = 0; g.W = q_len; g.s = 1; g.P = g.W < 33 ? 33 : g.W;
	}
	// lane 31 -> lane 0 hand-over: rows of up to NW_RING columns go through a (shared-memory) ring indexed j & rmask,
	// wider rows through a buffer indexed by j
	g.rmask = g.W <= NW_RING ? NW_RING - 1 : 0x7fffffff;
	g.d8 = pen.d8;
	g.R = (t_len + 31) >> 5;
	g.Tmax = g.s * ((t_len - 1) & 31) + g.P * (g.R - 1) + g.W;
	g.NEG = (t_len + q_len) * (pen.MM + pen.U + pen.W1);
	// rows of up to 32 * NW_RS_MAXC cells are swept row by row (nw_rs_*), wider ones run as the wavefront
	g.C = 0;
	if (rs && pen.d8 && g.W <= 32 * (rs < NW_RS_MAXC ? rs : NW_RS_MAXC) && (g.banded ? NW_RS_BAND : NW_RS_FULL)) {
		int c = (g.W + 31) >> 5;
		if (c == 5) c = 6;
		if (c == 7) c = 8;
		g.C = c;
	}
	return true;
}

NW_HD int nw_nuc(const uint64_t *seq, int pos) {
#if defined(__CUDA_ARCH__)
	return (int)((__ldg(seq + (pos >> 5)) << ((pos & 31) << 1)) >> 62);
#else
	return (int)((seq[pos >> 5] << ((pos & 31) << 1)) >> 62);
#endif
}

struct NwLane {
	int i, u;            // row and in-row offset of the NEXT step
	int ua, ub;          // in-row offsets of the first / last cell of row i (ua > ub: no cell)
	int joff;            // column of the cell at offset u is u + joff
	int flags;           // per-row facts, NWF_*
	int Dleft0, Ddiag0;  // what lies right of / diagonal to the first cell when the row starts at the matrix border
	int tn;              // template base of row i
	unsigned long long drow;   // d[tn][0..4] packed as signed bytes (when pen.d8)
	int act, qn;         // the next step computes a cell; its query base
	int myD, myP;        // last cell computed (what the lane below receives)
	int Dleft, Qleft;    // D, Q of (i, j-1)
	int Ddiag;           // D of (i-1, j-1)
	int nxD, nxP;        // lane 0: prefetched (D,P) of (i-1, j) for the next step
	int colBest, colBestI;   // k < 0: best D over the rows' last column (query start), first maximum in fill order
};

#define NWF_EDGE 1      // banded: the last cell of the row is the band's left edge (no vertical move)
#define NWF_TRACK 2     // k < 0 and the row reaches the query start: its last cell competes for the start cell
#define NWF_LAST 4      // last row (m = 0): D goes to lastD
#define NWF_BORDER 8    // the row's first cell touches the boundary column
#define NWF_ROW0 16     // first row: the row above is the analytic boundary row

NW_HD unsigned long long nw_pack_row(const NwPen &pen, int tn) {
	unsigned long long r = 0;
	for (int b = 0; b < 5; ++b) r |= (unsigned long long)(unsigned char)pen.d[tn * 5 + b] << (8 * b);
	return r;
}

// per-row constants of row L.i (called once per row, every P steps)
NW_HD void nw_row_setup(const NwGeo &g, const NwPen &pen, NwLane &L, const uint64_t *tseq, int t_s) {
	const int i = L.i;
	if (i >= g.t_len) { L.ua = 1; L.ub = 0; L.flags = 0; return; }
	const int jl = g.jlo(i), jh = g.jhi(i);
	L.joff = g.off(i);
	L.ua = jl - L.joff; L.ub = jh - L.joff;
	L.tn = nw_nuc(tseq, t_s + g.t_len - 1 - i);
	if (g.d8) L.drow = nw_pack_row(pen, L.tn);
	L.flags = (g.banded ? NWF_EDGE : 0) | ((g.k < 0 && jh == g.q_len - 1) ? NWF_TRACK : 0) | (i == g.t_len - 1 ? NWF_LAST : 0) |
	          (jl == 0 ? NWF_BORDER : 0) | (i == 0 ? NWF_ROW0 : 0);
	L.Dleft0 = jl == 0 ? g.bcol(i) : g.NEG;   // outside the band: nw.c:1031-1034
	L.Ddiag0 = g.bcol(i - 1);
}

// qlast points at the LAST query base of the window: the cell at column j reads qlast[-j]
NW_HD void nw_lane_arm(NwLane &L, const uint8_t *qlast) {
	L.act = L.u >= L.ua && L.u <= L.ub;
	L.qn = L.act ? qlast[-(L.u + L.joff)] : 0;
}

NW_HD void nw_lane_init(const NwGeo &g, const NwPen &pen, NwLane &L, int lane, const uint64_t *tseq, int t_s, const uint8_t *q) {
	L.i = lane; L.u = -g.s * lane;
	L.joff = 0; L.tn = 0; L.drow = 0; L.Dleft0 = 0; L.Ddiag0 = 0;
	nw_row_setup(g, pen, L, tseq, t_s);
	nw_lane_arm(L, q + g.q_len - 1);
	L.myD = 0; L.myP = g.NEG; L.Dleft = 0; L.Qleft = g.NEG; L.Ddiag = 0; L.nxD = 0; L.nxP = g.NEG;
	L.colBest = g.NEG; L.colBestI = 0x7fffffff;
}

// One step of one lane. aD/aP: (D,P) the lane above produced in the previous step (ignored by lane 0).
// qlast points at the last query base of the window; Estep at the 32 traceback bytes of this step. `hand` is the
// lane 31 -> lane 0 hand-over buffer (ring or row buffer, indexed j & g.rmask); the last row's D goes to lastD.
NW_HD void nw_lane_step(const NwGeo &g, const NwPen &pen, NwLane &L, const int lane, int aD, int aP,
                        const uint64_t *tseq, const int t_s, const uint8_t *qlast, uint8_t *Estep, NwRow *hand, int *lastD) {
	if (L.act) {
		const int u = L.u, j = u + L.joff;
		if (lane == 0) {
			if (L.flags & NWF_ROW0) { aD = g.brow(j); aP = g.NEG; }
			else { aD = L.nxD; aP = L.nxP; }
		}
		if (u == L.ua) {   // first cell of the row: what lies to its right and on its diagonal
			L.Dleft = L.Dleft0; L.Qleft = g.NEG;
			if (L.flags & NWF_BORDER) L.Ddiag = L.Ddiag0;
			else if (lane == 0) L.Ddiag = (L.flags & NWF_ROW0) ? g.brow(j - 1) : hand[(j - 1) & g.rmask].D;
		}
		const int sub = g.d8 ? (int)(signed char)(L.drow >> (L.qn << 3)) : pen.d[L.tn * 5 + L.qn];
		const int W1 = g.W1, U = g.U;
		const bool last = u == L.ub;
		int D, Q, P, e, fl = 0, x;
		if ((L.flags & NWF_EDGE) && last) {   // left edge of the band: no vertical move (nw.c:1076-1102)
			Q = L.Dleft + W1;
			x = L.Qleft + U;
			if (Q < x) { Q = x; e = 3; } else { e = 2; fl = 16; }
			P = g.NEG;
			D = L.Ddiag + sub;
			if (Q <= D) e = 1; else D = Q;
		} else {                              // nw.c:166-212
			Q = L.Dleft + W1; P = aD + W1;
			if (Q < P) { D = P; e = 4; } else { D = Q; e = 2; }
			x = L.Qleft + U;
			if (Q < x) { Q = x; if (D <= x) { D = x; e = 3; } } else fl |= 16;
			x = aP + U;
			if (P < x) { P = x; if (D <= x) { D = x; e = 5; } } else fl |= 32;
			x = L.Ddiag + sub;
			if (D <= x) { D = x; e = 1; }
		}
		Estep[lane] = (uint8_t)(fl | e);
		L.Dleft = D; L.Qleft = Q; L.myD = D; L.myP = P;
		if (last && (L.flags & NWF_TRACK) && L.colBest < D) { L.colBest = D; L.colBestI = L.i; }
		if (L.flags & NWF_LAST) lastD[j] = D;
		else if (lane == 31) { NwRow o; o.D = D; o.P = P; hand[j & g.rmask] = o; }
	}
	L.Ddiag = aD;
	if (++L.u == g.P) {
		L.u = 0; L.i += 32;
		nw_row_setup(g, pen, L, tseq, t_s);
	}
	nw_lane_arm(L, qlast);
}

// lane 0, after the step's stores are visible: fetch what the next step needs from the hand-over buffer
NW_HD void nw_lane0_prefetch(const NwGeo &g, NwLane &L, const NwRow *hand) {
	if (L.act && !(L.flags & NWF_ROW0)) { const NwRow v = hand[(L.u + L.joff) & g.rmask]; L.nxD = v.D; L.nxP = v.P; }
}

// ------------------------------------------------------------------------------------------------ row sweep
// Rows of up to 32 * NW_RS_MAXC cells (every default band, every short gap between two MEMs) are filled row by row:
// lane l owns the C consecutive in-row offsets u = l*C .. l*C+C-1 of every row and keeps (D, P) of the previous row's
// own cells in registers. The vertical and diagonal inputs are then the lane's own registers (one neighbour value
// per row comes by shuffle); the horizontal run Q(u) = max(D(u-1) + W1, Q(u-1) + U) is a max-plus prefix:
//   pass 1  per cell: P (open / extend, flag 32), diag + substitution; H = max(P, diag) is what the cell is worth
//           without its horizontal run; the lane folds its cells into A = the run leaving the chunk;
//   scan    5 shuffle rounds of a prefix maximum over the lanes give every lane the exact Q of its first cell,
//           because D(u-1) = max(H(u-1), Q(u-1)) makes Q(u) = max(H(u-1) + W1, Q(u-1) + max(W1, U));
//   pass 2  per cell, sequentially inside the lane: Q open / extend (flag 16), the reference's tie order among
//           Q, P and the diagonal, the traceback byte. The first cell's open / extend flag needs (D, Q) of the left
//           lane's last cell: two shuffles after the pass and a fix of that byte.
// The tie order of nw.c:166-212 in closed form: the vertical run wins against the horizontal one iff
// P >= Q + [P opened here]; the winner's code is 5 - [P opened] or 3 - [Q opened]; the diagonal wins iff it is >=.
#define NW_NINF (-0x3f000000)   // below every value a valid cell can hold, far from overflow

template <int C> struct NwRsLane {
	int pD[C], pP[C];          // previous row's own cells; between the passes: diag + substitution and P of this row
	int qs[C];                 // 8 * query base of the own cells' columns
	int colBest, colBestI;     // as NwLane
};

struct NwRsRow {
	int ua, ub, border, track, gen, Dleft0, Ddiag0, Qstart;
	unsigned lo, hi;           // d[tn][0..3] as signed bytes, d[tn][4]
};

NW_HD unsigned nw_fsr(unsigned lo, unsigned hi, unsigned sh) {   // (hi:lo) >> sh, sh <= 32
#if defined(__CUDA_ARCH__)
	return __funnelshift_rc(lo, hi, sh);
#else
	return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (sh > 32 ? 32 : sh));
#endif
}

NW_HD int nw_max(int a, int b) { return a < b ? b : a; }

// the per-template-base substitution rows, 8 bytes each (5 entries), written once per call
NW_HD unsigned long long nw_rs_tab(const NwPen &pen, int tn) {
	unsigned long long r = 0;
	for (int b = 0; b < 4; ++b) r |= (unsigned long long)(unsigned char)pen.d[tn * 5 + b] << (8 * b);
	return r | (unsigned long long)(unsigned char)pen.d[tn * 5 + 4] << 32;
}

// query base (times 8) under in-row offset u of row i; 0 outside the query
template <bool BANDED> NW_HD int nw_rs_q8(const NwGeo &g, int i, int u, const uint8_t *qlast) {
	const int j = BANDED ? u + g.a + i : u;
	return (j >= 0 && j < g.q_len) ? (int)qlast[-j] << 3 : 0;
}

template <int C, bool BANDED> NW_HD void nw_rs_init(const NwGeo &g, NwRsLane<C> &L, int lane, const uint8_t *qlast) {
	for (int c = 0; c < C; ++c) {
		const int u = lane * C + c;
		const int j = BANDED ? u + g.a - 1 : u;   // column the cell of row -1 (the analytic boundary row) stands for
		L.pD[c] = g.brow(j); L.pP[c] = g.NEG;
		L.qs[c] = nw_rs_q8<BANDED>(g, 0, u, qlast);
	}
	L.colBest = g.NEG; L.colBestI = 0x7fffffff;
}

NW_HD void nw_rs_row(const NwGeo &g, NwRsRow &R, int i, unsigned long long tabrow) {
	const int jl = g.jlo(i), jh = g.jhi(i), off = g.off(i);
	R.ua = jl - off; R.ub = jh - off;
	R.border = jl == 0;
	R.gen = R.ua > 0;   // the row starts inside the lanes' range (first rows of a band): the general per-cell tests run
	R.track = g.k < 0 && jh == g.q_len - 1;
	R.Dleft0 = jl == 0 ? g.bcol(i) : g.NEG;
	R.Ddiag0 = g.bcol(i - 1);
	R.Qstart = nw_max(R.Dleft0 + g.W1, g.NEG + g.U);
	R.lo = (unsigned)tabrow; R.hi = (unsigned)(tabrow >> 32);
}

// nbD / nbP: banded -- (D, P) of the right lane's first cell; full -- D of the left lane's last cell (previous row).
// Returns A, the horizontal run leaving the chunk as far as the lane's own cells decide it. Cells right of the row's
// last one compute on whatever they hold: nothing flows from them into a valid cell.
template <int C, bool BANDED>
NW_HD int nw_rs_pass1(const NwGeo &g, NwRsLane<C> &L, int lane, const NwRsRow &R, int nbD, int nbP, int Ue, int *fP) {
	const int u0 = lane * C;
	int dprev = nbD, r = NW_NINF;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int c = 0; c < C; ++c) {
		const int u = u0 + c;
		int upD, upP, dg;
		if (BANDED) {
			dg = L.pD[c];
			upD = c + 1 < C ? L.pD[c + 1 < C ? c + 1 : c] : nbD;
			upP = c + 1 < C ? L.pP[c + 1 < C ? c + 1 : c] : nbP;
		} else {
			dg = dprev; upD = L.pD[c]; upP = L.pP[c]; dprev = upD;
		}
		if ((c == 0 || (BANDED && R.gen)) && u == R.ua) {   // the row's first cell
			if (R.border) dg = R.Ddiag0;
			r = R.Qstart;
		}
		dg += (int)(signed char)nw_fsr(R.lo, R.hi, (unsigned)L.qs[c]);
		const int Po = upD + g.W1, Pe = upP + g.U;
		const bool edge = BANDED && u == R.ub;   // the row's last cell has no vertical move (nw.c:1076-1102)
		fP[c] = (!edge && Po >= Pe) ? 1 : 0;
		const int P = edge ? g.NEG : nw_max(Po, Pe);
		L.pD[c] = dg; L.pP[c] = P;
		int H = nw_max(P, dg);
		if (BANDED && R.gen && u < R.ua) H = NW_NINF;
		r = nw_max(H + g.W1, r + Ue);
	}
	return r;
}

// Qfirst: the exact Q of the lane's first cell. e[]: traceback bytes (e[0] lacks the first cell's Q-opened flag).
template <int C, bool BANDED>
NW_HD void nw_rs_pass2(const NwGeo &g, NwRsLane<C> &L, int lane, const NwRsRow &R, int Qfirst, const int *fP, int *e, int *Dlast,
                       int *Qlast) {
	const int u0 = lane * C;
	int Dl = 0, Ql = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int c = 0; c < C; ++c) {
		const int u = u0 + c;
		int Q, cq, fq;
		if (c == 0) { Q = u == R.ua ? R.Qstart : Qfirst; cq = 3; fq = 0; }
		else {
			int Qo = Dl + g.W1, Qe = Ql + g.U;
			if (BANDED && R.gen && u == R.ua) { Qo = R.Dleft0 + g.W1; Qe = g.NEG + g.U; }
			const bool o = Qo >= Qe;
			Q = o ? Qo : Qe; cq = o ? 2 : 3; fq = o ? 16 : 0;
		}
		const int P = L.pP[c], dg = L.pD[c];
		const bool edge = BANDED && u == R.ub;
		const bool pw = !edge && P >= Q + fP[c];
		const int D1 = pw ? P : Q;
		int ec = pw ? 5 - fP[c] : cq;
		if (D1 <= dg) ec = 1;
		e[c] = ec + fq + (fP[c] << 5);
		const int D = nw_max(D1, dg);
		L.pD[c] = D;
		Dl = D; Ql = Q;
	}
	*Dlast = Dl; *Qlast = Ql;
}

// the first cell's Q-opened flag once (D, Q) of the cell left of it are known
template <int C> NW_HD int nw_rs_fix0(const NwGeo &g, int lane, const NwRsRow &R, int e0, int Dl, int Ql) {
	if (lane * C == R.ua) { Dl = R.Dleft0; Ql = g.NEG; }
	if (Dl + g.W1 >= Ql + g.U) {
		if ((e0 & 7) == 3) e0 = (e0 & ~7) | 2;
		e0 |= 16;
	}
	return e0;
}

// after the row: the row's last cell competes for the start cell (k < 0); next row's query bases
template <int C, bool BANDED>
NW_HD void nw_rs_rowend(const NwGeo &g, NwRsLane<C> &L, int lane, const NwRsRow &R, int i, int q8next) {
	if (R.track) {
		for (int c = 0; c < C; ++c)
			if (lane * C + c == R.ub && L.colBest < L.pD[c]) { L.colBest = L.pD[c]; L.colBestI = i; }
	}
	if (BANDED) {
		for (int c = 0; c + 1 < C; ++c) L.qs[c] = L.qs[c + 1];
		L.qs[C - 1] = q8next;
	}
}

// ---- interior rows of a band: a + i >= 1 and a + i + band <= q_len - 2. Such a row starts at in-row offset 0 away from the
// matrix border, ends at offset `band` (the edge cell, always the same lane and cell) and does not reach the query
// start: no border values, no partial rows, no start-cell tracking. Nearly every row of a long banded problem is one
// (all but ~band / 2 at either end), and what is left of the passes is about half the instructions of the general ones.
NW_HD void nw_rs_interior(const NwGeo &g, int *i1, int *i2) {
	int lo = 1 - g.a, hi = g.q_len - 1 - g.band - g.a;   // rows [lo, hi)
	if (lo < 0) lo = 0;
	if (hi > g.t_len) hi = g.t_len;
	if (!g.banded || hi < lo) hi = lo;
	*i1 = lo; *i2 = hi;
}

// ubc: which of the lane's cells is the row's edge cell (-1: none); lane0: the lane holds the row's first cell
template <int C>
NW_HD int nw_rs_pass1i(const NwGeo &g, NwRsLane<C> &L, bool lane0, int ubc, unsigned lo, unsigned hi, int nbD, int nbP, int Ue, int Qstart, int *fP) {
	int r = lane0 ? Qstart : NW_NINF;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int c = 0; c < C; ++c) {
		const int upD = c + 1 < C ? L.pD[c + 1 < C ? c + 1 : c] : nbD, upP = c + 1 < C ? L.pP[c + 1 < C ? c + 1 : c] : nbP;
		const int dg = L.pD[c] + (int)(signed char)nw_fsr(lo, hi, (unsigned)L.qs[c]);
		const int Po = upD + g.W1, Pe = upP + g.U;
		const bool edge = c == ubc;   // no vertical move into the band's last cell (nw.c:1076-1102)
		fP[c] = (!edge && Po >= Pe) ? 1 : 0;
		const int P = edge ? g.NEG : nw_max(Po, Pe);
		L.pD[c] = dg; L.pP[c] = P;
		r = nw_max(nw_max(P, dg) + g.W1, r + Ue);
	}
	return r;
}

template <int C>
NW_HD void nw_rs_pass2i(const NwGeo &g, NwRsLane<C> &L, bool lane0, int ubc, int Qfirst, int Qstart, const int *fP, int *e, int *Dlast, int *Qlast) {
	int Dl = 0, Ql = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int c = 0; c < C; ++c) {
		int Q, cq, fq;
		if (c == 0) { Q = lane0 ? Qstart : Qfirst; cq = 3; fq = 0; }
		else {
			const int Qo = Dl + g.W1, Qe = Ql + g.U;
			const bool o = Qo >= Qe;
			Q = o ? Qo : Qe; cq = o ? 2 : 3; fq = o ? 16 : 0;
		}
		const int P = L.pP[c], dg = L.pD[c];
		const bool pw = c != ubc && P >= Q + fP[c];
		const int D1 = pw ? P : Q;
		int ec = pw ? 5 - fP[c] : cq;
		if (D1 <= dg) ec = 1;
		e[c] = ec + fq + (fP[c] << 5);
		const int D = nw_max(D1, dg);
		L.pD[c] = D;
		Dl = D; Ql = Q;
	}
	*Dlast = Dl; *Qlast = Ql;
}

// the first cell's Q-opened flag; the row's first cell has nothing but the band's outside to its left
NW_HD int nw_rs_fix0i(const NwGeo &g, bool lane0, int e0, int Dl, int Ql) {
	if (lane0) { Dl = g.NEG; Ql = g.NEG; }
	if (Dl + g.W1 >= Ql + g.U) {
		if ((e0 & 7) == 3) e0 = (e0 & ~7) | 2;
		e0 |= 16;
	}
	return e0;
}

// the last row's D by query column (what nw_start_cell reads)
template <int C, bool BANDED> NW_HD void nw_rs_lastrow(const NwGeo &g, const NwRsLane<C> &L, int lane, int *lastD) {
	const int i = g.t_len - 1, off = g.off(i), ua = g.jlo(i) - off, ub = g.jhi(i) - off;
	for (int c = 0; c < C; ++c) {
		const int u = lane * C + c;
		if (u >= ua && u <= ub) lastD[u + off] = L.pD[c];
	}
}

// traceback byte at (m, qpos) in reference coordinates, including the analytic boundary row / column codes
NW_HD int nw_e_at(const NwGeo &g, const uint8_t *E, int m, int qpos) {
	if (m >= g.t_len) {   // row "all template consumed" (nw.c:118-151, 1002-1019)
		if (g.k == 2 || qpos >= g.q_len) return 0;
		return qpos == g.q_len - 1 ? 18 : 3;
	}
	const int i = g.t_len - 1 - m, j = g.q_len - 1 - qpos, jl = g.jlo(i);
	if (j < jl) {         // right of the row's last cell
		if (g.banded) return jl > 0 ? 37 : (0 < g.k ? 0 : 37);
		return 0 < g.k ? 0 : (m == g.t_len - 1 ? 36 : 5);
	}
	if (j > g.jhi(i)) return 0;   // unreachable by construction
	return E[g.eaddr(i, j)];
}

// traceback statistics (nw.c:850-887, 1143-1185) from the start cell
NW_HD void nw_walk(const NwGeo &g, const uint8_t *E, int m, int qp, NwStat &s) {
	int e;
	s.len = s.match = s.tGaps = s.qGaps = 0;
	while ((e = nw_e_at(g, E, m, qp)) != 0) {
		const int c = e & 7;
		if (c == 1) { ++s.match; ++m; ++qp; }
		else if (c >= 4) {
			while (!(nw_e_at(g, E, m, qp) >> 4)) { ++m; ++s.len; ++s.qGaps; }
			++s.qGaps; ++m;
		} else {
			while (!(nw_e_at(g, E, m, qp) >> 3)) { ++qp; ++s.len; ++s.tGaps; }
			++s.tGaps; ++qp;
		}
		++s.len;
	}
}

// start cell (nw.c:831-848, 1127-1141) given the reduced column maximum. rowBest/rowBestQ: k == -2 only, the LAST
// maximum of D along row m = 0 over the query positions nw_row0_range() names.
NW_HD void nw_row0_range(const NwGeo &g, int *qlo, int *qhi) {
	const int last = g.t_len - 1;
	const int lo0 = g.q_len - 1 - g.jhi(last), hi0 = g.q_len - 1 - g.jlo(last);
	*qlo = lo0; *qhi = hi0;
	if (g.banded) {
		// the reference scans its row buffer n = en..bq-1 (nw.c:1135); en = rows whose band start was clamped
		int en = 0;
		if (lo0 == 0) { en = -(g.a + last + g.band - (g.q_len - 1)); if (en < 0) en = 0; }
		const int qe = lo0 + (g.band - en);
		if (qe < *qhi) *qhi = qe;
	}
}

NW_HD void nw_start_cell(const NwGeo &g, int colBest, int colBestI, const int *lastD, int rowBest, int rowBestQ,
                         int *best_m, int *best_q, int *score) {
	const int last = g.t_len - 1;
	const int lo0 = g.q_len - 1 - g.jhi(last);
	if (g.k < 0) {
		if (colBestI == 0x7fffffff) { *best_m = 0; *score = g.NEG; } else { *best_m = g.t_len - 1 - colBestI; *score = colBest; }
		*best_q = 0;
		if (g.banded && *best_m == 0) { *best_q = lo0; *score = lastD[g.q_len - 1 - lo0]; }
		if (g.k == -2 && rowBestQ >= 0 && *score <= rowBest) { *score = rowBest; *best_m = 0; *best_q = rowBestQ; }
	} else {
		*best_m = 0; *best_q = g.banded ? lo0 : 0;
		*score = lastD[g.q_len - 1 - *best_q];
	}
}

NW_HD bool nw_trivial(const NwPen &pen, int t_len, int q_len, NwStat &s) {   // nw.c:663-684
	if (t_len != 0 && q_len != 0) return false;
	s.score = s.len = s.pos = s.match = s.tGaps = s.qGaps = 0;
	if (t_len != q_len) {
		if (t_len == 0) { s.len = q_len; s.tGaps = q_len; s.score = pen.W1 + (q_len - 1) * pen.U; }
		else { s.len = t_len; s.qGaps = t_len; s.score = pen.W1 + (t_len - 1) * pen.U; }
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ one thread per problem
// The short tails and gaps of short reads (C1/C2: ~130 cells per read, a few bases by a few bases) are far too small
// for a warp: NW_score (nw.c:642-890) of a FULL matrix filled by ONE thread, row by row, 32 independent problems per
// warp. The previous row (D, P) lives in `rows` (q_len entries, element x at rows[x * rstride]: shared memory,
// transposed so that the lanes of a warp never conflict), the traceback bytes in `E` (cell c at E[c * estride]: the
// warp's lanes write neighbouring bytes of one sector as long as they are at the same cell). Same recurrence, tie
// rules, start cell and walk as nw_lane_step / nw_start_cell / nw_walk; checked against the oracle by the emulator.
NW_HD int nw_thread_e_at(const uint8_t *E, size_t estride, int t_len, int q_len, int k, int m, int qpos) {
	if (m >= t_len) {
		if (k == 2 || qpos >= q_len) return 0;
		return qpos == q_len - 1 ? 18 : 3;
	}
	if (qpos >= q_len) return 0 < k ? 0 : (m == t_len - 1 ? 36 : 5);
	return E[((size_t)(t_len - 1 - m) * (size_t)q_len + (size_t)(q_len - 1 - qpos)) * estride];
}

// qs: the window's query bases in FILL order (qs[j * qstride] = 8 * q[q_len - 1 - j]: the shift that selects the
// substitution byte), staged by the caller next to `rows`. D8: the substitution rows fit signed bytes (pen.d8) and
// tab[tn] = nw_pack_row(pen, tn).
template <bool D8>
NW_HD void nw_thread(const NwPen &pen, const unsigned long long *tab, const uint64_t *tseq, int t_s, int t_len, const uint8_t *qs,
                     int qstride, int q_len, int k, NwRow *rows, int rstride, uint8_t *E, size_t estride, NwStat *out) {
	const int W1 = pen.W1, U = pen.U;
	const int NEG = (t_len + q_len) * (pen.MM + U + W1);
	for (int j = 0; j < q_len; ++j) { NwRow r; r.D = k == 2 ? 0 : W1 + j * U; r.P = NEG; rows[(size_t)j * rstride] = r; }
	int colBest = NEG, colBestI = 0x7fffffff;
	// ONE loop over the cells in fill order (row by row): the 32 problems of a warp stay in lock step however their
	// shapes differ, and the warp takes max(cells) iterations instead of max(t_len) * max(q_len). The cell is the
	// recurrence of nw.c:166-212 in closed form (see the row sweep above): the vertical run wins against the horizontal
	// one iff P >= Q + [P opened here]; the winner's code is 5 - [P opened] or 3 - [Q opened]; the diagonal wins ties.
	uint8_t *e = E;
	NwRow *rp = rows;
	const uint8_t *qp = qs;
	int i = 0, j = 0, tn = 0, Dleft = 0, Qleft = 0, Ddiag = 0;
	unsigned long long drow = 0;
	for (int c = t_len * q_len; c > 0; --c, e += estride) {
		if (j == 0) {   // a new row: template base, what lies beside and diagonal to its first cell
			tn = nw_nuc(tseq, t_s + t_len - 1 - i);
			if (D8) drow = tab[tn];
			Dleft = 0 < k ? 0 : W1 + i * U; Qleft = NEG;
			Ddiag = i == 0 ? 0 : (0 < k ? 0 : W1 + (i - 1) * U);
			rp = rows; qp = qs;
		}
		const int q8 = *qp;
		const NwRow a = *rp;
		const int sub = D8 ? (int)(signed char)(drow >> q8) : pen.d[tn * 5 + (q8 >> 3)];
		const int Qo = Dleft + W1, Qe = Qleft + U, Po = a.D + W1, Pe = a.P + U, dg = Ddiag + sub;
		const bool qo = Qo >= Qe, po = Po >= Pe;
		const int Q = qo ? Qo : Qe, P = po ? Po : Pe;
		const bool pw = P - (po ? 1 : 0) >= Q;
		const int D1 = pw ? P : Q;
		int ec = pw ? (po ? 4 : 5) : (qo ? 2 : 3);
		const bool dw = D1 <= dg;
		const int D = dw ? dg : D1;
		if (dw) ec = 1;
		*e = (uint8_t)(ec | (qo ? 16 : 0) | (po ? 32 : 0));
		NwRow o; o.D = D; o.P = P;
		*rp = o;
		rp += rstride; qp += qstride;
		Ddiag = a.D; Dleft = D; Qleft = Q;
		if (++j == q_len) {
			if (k < 0 && colBest < D) { colBest = D; colBestI = i; }   // the row's cell at the query start competes (nw.c:217-220)
			j = 0; ++i;
		}
	}
	int best_m = 0, best_q = 0, score;
	if (k < 0) {
		if (colBestI == 0x7fffffff) score = NEG; else { best_m = t_len - 1 - colBestI; score = colBest; }
		if (k == -2) {   // last maximum along row m = 0 (nw.c:235-243)
			int rb = NEG, rq = -1;
			for (int qp2 = 0; qp2 < q_len; ++qp2) { const int v = rows[(size_t)(q_len - 1 - qp2) * rstride].D; if (rq < 0 || v >= rb) { rb = v; rq = qp2; } }
			if (rq >= 0 && score <= rb) { score = rb; best_m = 0; best_q = rq; }
		}
	} else score = rows[(size_t)(q_len - 1) * rstride].D;
	NwStat s;
	s.len = s.match = s.tGaps = s.qGaps = 0;
	int m = best_m, qpos = best_q, c;
	while ((c = nw_thread_e_at(E, estride, t_len, q_len, k, m, qpos)) != 0) {   // nw.c:850-887
		const int d = c & 7;
		if (d == 1) { ++s.match; ++m; ++qpos; }
		else if (d >= 4) {
			while (!(nw_thread_e_at(E, estride, t_len, q_len, k, m, qpos) >> 4)) { ++m; ++s.len; ++s.qGaps; }
			++s.qGaps; ++m;
		} else {
			while (!(nw_thread_e_at(E, estride, t_len, q_len, k, m, qpos) >> 3)) { ++qpos; ++s.len; ++s.tGaps; }
			++s.tGaps; ++qpos;
		}
		++s.len;
	}
	s.score = score; s.pos = 0;
	*out = s;
}

#if defined(__CUDACC__)
// per-warp scratch
struct NwScratch {
	NwRow *ring;      // NW_RING entries in shared memory: lane 31 -> lane 0 hand-over of rows up to NW_RING wide
	NwRow *rowbuf;    // global, q_cap entries: the same for wider rows; lastD (q_cap ints) and E (e_cap bytes) follow it
	size_t e_cap;
	int q_cap;
#ifdef NW_SCRATCH_WIDE
	int *lastD_; uint8_t *E_;
	__device__ __forceinline__ int *lastD() const { return lastD_; }
	__device__ __forceinline__ uint8_t *E() const { return E_; }
	__device__ __forceinline__ void finish() { lastD_ = (int *)(rowbuf + q_cap); E_ = (uint8_t *)(rowbuf + q_cap) + 4 * (size_t)q_cap; }
#else
	__device__ __forceinline__ int *lastD() const { return (int *)(rowbuf + q_cap); }
	__device__ __forceinline__ uint8_t *E() const { return (uint8_t *)(rowbuf + q_cap) + 4 * (size_t)q_cap; }
	__device__ __forceinline__ void finish() {}
#endif
};

// status of nw_warp
#define NW_OK 0
#define NW_TOO_BIG 1      // scratch too small: the caller re-runs the problem on the large-scratch path
#define NW_BAD_BAND 2

// aligned rows of one call (NW / NW_band, nw.c:250-305): template, match and query row, written from column 0.
// tseq/t_s/q give the bases; all null for the score-only calls.
struct NwRows { uint8_t *t, *s, *q; };

// The walk of nw_walk with up to 32 cells examined per memory round trip: the lanes read the traceback bytes along
// the direction the path is moving (diagonal, down a P run, along a Q run) and a ballot finds where the run ends.
__device__ __forceinline__ void nw_walk_warp(const NwGeo &g, const uint8_t *E, int m, int qp, NwStat &s, const NwRows *rows,
                                             const uint64_t *tseq, int t_s, const uint8_t *q) {
	const int lane = threadIdx.x & 31;
	s.len = s.match = s.tGaps = s.qGaps = 0;
	for (;;) {
		const int e = nw_e_at(g, E, m + lane, qp + lane);
		const unsigned nd = __ballot_sync(0xffffffffu, (e & 7) != 1);
		const int r = nd ? __ffs(nd) - 1 : 32;   // leading diagonal steps
		if (rows && lane < r) {
			const uint8_t tb = (uint8_t)nw_nuc(tseq, t_s + m + lane), qb = q[qp + lane];
			rows->t[s.len + lane] = tb; rows->q[s.len + lane] = qb; rows->s[s.len + lane] = tb == qb ? '|' : '_';
		}
		s.match += r; s.len += r; m += r; qp += r;
		if (r == 32) continue;
		const int c = __shfl_sync(0xffffffffu, e, r);
		if (c == 0) break;
		if ((c & 7) >= 4) {   // P run: down the column until a cell whose run opens there
			for (;;) {
				const int f = nw_e_at(g, E, m + lane, qp);
				const unsigned stop = __ballot_sync(0xffffffffu, (f >> 4) != 0);
				const int n = stop ? __ffs(stop) - 1 : 32;
				const int w = stop ? n + 1 : 32;   // columns written: the run so far and, at its end, the closing cell
				if (rows && lane < w) { rows->t[s.len + lane] = (uint8_t)nw_nuc(tseq, t_s + m + lane); rows->q[s.len + lane] = 5; rows->s[s.len + lane] = '_'; }
				m += n; s.len += n; s.qGaps += n;
				if (stop) break;
			}
			++s.qGaps; ++m;
		} else {              // Q run: along the row
			for (;;) {
				const int f = nw_e_at(g, E, m, qp + lane);
				const unsigned stop = __ballot_sync(0xffffffffu, (f >> 3) != 0);
				const int n = stop ? __ffs(stop) - 1 : 32;
				const int w = stop ? n + 1 : 32;
				if (rows && lane < w) { rows->t[s.len + lane] = 5; rows->q[s.len + lane] = q[qp + lane]; rows->s[s.len + lane] = '_'; }
				qp += n; s.len += n; s.tGaps += n;
				if (stop) break;
			}
			++s.tGaps; ++qp;
		}
		++s.len;
	}
}

// Row sweep of one problem: fills E (row-major, 32*C bytes per row) and lastD, leaves the lane's column maximum.
template <int C, bool BANDED>
__device__ __noinline__ void nw_rs_fill(const NwGeo &gin, const uint64_t *__restrict__ tseq, int t_s, const uint8_t *q,
                                        const unsigned long long *tab, uint8_t *E, int *lastD, int *cbo, int *cio) {
	const NwGeo g = gin;   // by value: the traceback stores below must not force reloads of the geometry
	const int lane = threadIdx.x & 31;
	const unsigned full = 0xffffffffu;
	const uint8_t *qlast = q + g.q_len - 1;
	NwRsLane<C> L;
	nw_rs_init<C, BANDED>(g, L, lane, qlast);
	const int Ue = g.W1 > g.U ? g.W1 : g.U, lstep = lane * C * Ue;
	uint8_t *Erow = E + lane * C;
	int tpos = t_s + g.t_len - 1;
	int i1 = 0, i2 = 0;   // the band's interior rows
	if (BANDED) nw_rs_interior(g, &i1, &i2);
	const bool lane0 = lane == 0;
	const int ubc = (g.band >= lane * C && g.band < lane * C + C) ? g.band - lane * C : -1;
	const int Qs_in = nw_max(g.NEG + g.W1, g.NEG + g.U);
	auto store_row = [&](const int *e) {
		if (C == 1) Erow[0] = (uint8_t)e[0];
		else if (C == 2) *(uint16_t *)Erow = (uint16_t)(e[0] | e[1 % C] << 8);
		else if (C == 3) { Erow[0] = (uint8_t)e[0]; Erow[1] = (uint8_t)e[1 % C]; Erow[2] = (uint8_t)e[2 % C]; }
		else if (C == 4) *(uint32_t *)Erow = (uint32_t)(e[0] | e[1 % C] << 8 | e[2 % C] << 16 | e[3 % C] << 24);
		else if (C == 6) {
			((uint16_t *)Erow)[0] = (uint16_t)(e[0] | e[1 % C] << 8);
			((uint16_t *)Erow)[1] = (uint16_t)(e[2 % C] | e[3 % C] << 8);
			((uint16_t *)Erow)[2] = (uint16_t)(e[4 % C] | e[5 % C] << 8);
		} else {
			uint2 w;
			w.x = (uint32_t)(e[0] | e[1 % C] << 8 | e[2 % C] << 16 | e[3 % C] << 24);
			w.y = (uint32_t)(e[4 % C] | e[5 % C] << 8 | e[6 % C] << 16 | e[7 % C] << 24);
			*(uint2 *)Erow = w;
		}
	};
	for (int i = 0; i < g.t_len; ++i, Erow += 32 * C, --tpos) {
		if (BANDED && i >= i1 && i < i2) {   // interior row: the lean passes
			const unsigned long long tr = tab[nw_nuc(tseq, tpos)];
			const int q8n = nw_rs_q8<true>(g, i + 1, lane * C + C - 1, qlast);
			const int nbD = __shfl_down_sync(full, L.pD[0], 1), nbP = __shfl_down_sync(full, L.pP[0], 1);
			int fP[C];
			int B = nw_rs_pass1i<C>(g, L, lane0, ubc, (unsigned)tr, (unsigned)(tr >> 32), nbD, nbP, Ue, Qs_in, fP) - lstep;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) B = max(B, __shfl_up_sync(full, B, o));
			const int Qf = __shfl_up_sync(full, B + lstep, 1);
			int e[C], Dlast, Qlast;
			nw_rs_pass2i<C>(g, L, lane0, ubc, Qf, Qs_in, fP, e, &Dlast, &Qlast);
			const int Dl = __shfl_up_sync(full, Dlast, 1), Ql = __shfl_up_sync(full, Qlast, 1);
			e[0] = nw_rs_fix0i(g, lane0, e[0], Dl, Ql);
			store_row(e);
#pragma unroll
			for (int c = 0; c + 1 < C; ++c) L.qs[c] = L.qs[c + 1];
			L.qs[C - 1] = q8n;
			continue;
		}
		NwRsRow R;
		nw_rs_row(g, R, i, tab[nw_nuc(tseq, tpos)]);
		int q8n = 0;
		if (BANDED) q8n = nw_rs_q8<true>(g, i + 1, lane * C + C - 1, qlast);   // next row's new column, used at the row's end
		int nbD, nbP = 0;
		if (BANDED) { nbD = __shfl_down_sync(full, L.pD[0], 1); nbP = __shfl_down_sync(full, L.pP[0], 1); }
		else nbD = __shfl_up_sync(full, L.pD[C - 1], 1);
		int fP[C];
		int B = nw_rs_pass1<C, BANDED>(g, L, lane, R, nbD, nbP, Ue, fP) - lstep;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) B = max(B, __shfl_up_sync(full, B, o));   // lanes below o get their own value back
		int Qf = __shfl_up_sync(full, B + lstep, 1);
		if (lane == 0) Qf = NW_NINF;
		int e[C], Dlast, Qlast;
		nw_rs_pass2<C, BANDED>(g, L, lane, R, Qf, fP, e, &Dlast, &Qlast);
		const int Dl = __shfl_up_sync(full, Dlast, 1), Ql = __shfl_up_sync(full, Qlast, 1);
		e[0] = nw_rs_fix0<C>(g, lane, R, e[0], Dl, Ql);
		store_row(e);
		nw_rs_rowend<C, BANDED>(g, L, lane, R, i, q8n);
	}
	nw_rs_lastrow<C, BANDED>(g, L, lane, lastD);
	*cbo = L.colBest; *cio = L.colBestI;
}

template <bool BANDED, int MAXC>
__device__ __forceinline__ void nw_rs_dispatch(const NwGeo &g, const uint64_t *__restrict__ tseq, int t_s, const uint8_t *q,
                                               const unsigned long long *tab, uint8_t *E, int *lastD, int *cb, int *ci) {
	if (BANDED ? !NW_RS_BAND : !NW_RS_FULL) return;
	// only the widths the kernel was built for are reachable from it: its register count is the widest sweep's
	if (g.C == 1) nw_rs_fill<1, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
	if constexpr (MAXC >= 2) if (g.C == 2) nw_rs_fill<2, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
	if constexpr (MAXC >= 3) if (g.C == 3) nw_rs_fill<3, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
	if constexpr (MAXC >= 4) if (g.C == 4) nw_rs_fill<4, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
	if constexpr (MAXC >= 6) if (g.C == 6) nw_rs_fill<6, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
	if constexpr (MAXC >= 8) if (g.C == 8) nw_rs_fill<8, BANDED>(g, tseq, t_s, q, tab, E, lastD, cb, ci);
}

// How the out-of-line nw_warp takes the scratch descriptor: by reference it lives in the caller's local memory, by value
// in registers across the call. The short-read build of the pair kernel gains 6 % by value, the long-read build loses
// 4-6 % (profiles/r01_ab_nwscratch_byvalue.log), so the translation unit chooses.
#ifdef NW_SCRATCH_BYVAL
typedef const NwScratch NwScratchArg;
#else
typedef const NwScratch &NwScratchArg;
#endif

// All 32 lanes call with identical arguments; every lane returns the same result. RS: rows of up to 32 * RS cells run
// as the row sweep (0: none).
template <int RS>
__device__ __noinline__ int nw_warp(const NwPen &pen, const uint64_t *__restrict__ tseq, const uint8_t *query, int k, int t_s,
                                    int t_e, int q_s, int q_e, int band, NwScratchArg ws, NwStat *out,
                                    unsigned long long *cells, const NwRows *rows = nullptr) {
	const int lane = threadIdx.x & 31;
	const int t_len = t_e - t_s, q_len = q_e - q_s;
	NwStat s;
	if (nw_trivial(pen, t_len, q_len, s)) {   // nw.c:49-85: one side empty -> all-gap rows
		if (rows) {
			if (t_len == 0) for (int i = lane; i < q_len; i += 32) { rows->t[i] = 5; rows->s[i] = '_'; rows->q[i] = query[q_s + i]; }
			else for (int i = lane; i < t_len; i += 32) { rows->t[i] = (uint8_t)nw_nuc(tseq, t_s + i); rows->s[i] = '_'; rows->q[i] = 5; }
			__syncwarp();
		}
		*out = s;
		return NW_OK;
	}
#ifndef KG_NO_FAST11
	if (t_len == 1 && q_len == 1 && k == 0 && !rows) {   // the single mismatch between two MEMs: one cell, closed form
		const int W1 = pen.W1, U = pen.U, NEG = 2 * (pen.MM + U + W1);
		const int sub = pen.d[nw_nuc(tseq, t_s) * 5 + query[q_s]];
		// the cell of nw.c:166-212 with Dleft = D(0,-1) = W1, aD = D(-1,0) = W1, Qleft = aP = NEG, Ddiag = D(-1,-1) = 0
		int Q = W1 + W1, P = W1 + W1, D, e, fl = 0, x;
		if (Q < P) { D = P; e = 4; } else { D = Q; e = 2; }
		x = NEG + U;
		if (Q < x) { Q = x; if (D <= x) { D = x; e = 3; } } else fl |= 16;
		if (P < x) { P = x; if (D <= x) { D = x; e = 5; } } else fl |= 32;
		x = sub;
		if (D <= x) { D = x; e = 1; }
		if (e == 1 || fl) {
			// diagonal: one column. Otherwise the run closes in this cell (flag set) and the walk crosses one gap of
			// each kind through the boundary codes (36 / 18): two columns
			s.score = D; s.pos = 0;
			if (e == 1) { s.len = 1; s.match = 1; s.tGaps = 0; s.qGaps = 0; }
			else { s.len = 2; s.match = 0; s.tGaps = 1; s.qGaps = 1; }
			if (cells) *cells += 1;
			*out = s;
			return NW_OK;
		}
	}
#endif
	NwGeo g;
	if (!nw_geo_init(g, pen, t_len, q_len, k, band, RS)) return NW_BAD_BAND;
	if (g.ebytes() > ws.e_cap || q_len + 1 > ws.q_cap) return NW_TOO_BIG;
	if (cells) *cells += (unsigned long long)t_len * (unsigned long long)(g.banded ? g.band + 1 : q_len);
	const uint8_t *q = query + q_s;
	const uint8_t *E = ws.E();
	int cb, ci;
	if (RS && g.C) {   // row sweep; the substitution rows go to the (otherwise unused) shared-memory ring
		unsigned long long *tab = (unsigned long long *)ws.ring;
		__syncwarp();
		if (lane < 5) tab[lane] = nw_rs_tab(pen, lane);
		__syncwarp();
		uint8_t *E8 = (uint8_t *)(((uintptr_t)ws.E() + 7) & ~(uintptr_t)7);
		if (g.banded) nw_rs_dispatch<true, RS>(g, tseq, t_s, q, tab, E8, ws.lastD(), &cb, &ci);
		else nw_rs_dispatch<false, RS>(g, tseq, t_s, q, tab, E8, ws.lastD(), &cb, &ci);
		E = E8;
	} else {
		NwRow *hand = g.rmask == NW_RING - 1 ? ws.ring : ws.rowbuf;
		NwLane L;
		nw_lane_init(g, pen, L, lane, tseq, t_s, q);
		const uint8_t *qlast = q + q_len - 1;
		uint8_t *Estep = ws.E();
		int *lastD = ws.lastD();
		for (int T = g.Tmax; T > 0; --T, Estep += 32) {
			const int aD = __shfl_up_sync(0xffffffffu, L.myD, 1), aP = __shfl_up_sync(0xffffffffu, L.myP, 1);
			nw_lane_step(g, pen, L, lane, aD, aP, tseq, t_s, qlast, Estep, hand, lastD);
			__syncwarp();
			if (lane == 0) nw_lane0_prefetch(g, L, hand);
		}
		cb = L.colBest; ci = L.colBestI;
	}
	__syncwarp();
	// column maximum: largest D, ties -> smallest i (first in fill order)
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		const int ov = __shfl_xor_sync(0xffffffffu, cb, o), oi = __shfl_xor_sync(0xffffffffu, ci, o);
		if (ov > cb || (ov == cb && oi < ci)) { cb = ov; ci = oi; }
	}
	int rb = g.NEG, rq = -1;
	if (k == -2) {   // last maximum along row m = 0
		int qlo, qhi;
		nw_row0_range(g, &qlo, &qhi);
		for (int qp = qlo + lane; qp <= qhi; qp += 32) {
			const int v = ws.lastD()[q_len - 1 - qp];
			if (rq < 0 || v >= rb) { rb = v; rq = qp; }
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			const int ov = __shfl_xor_sync(0xffffffffu, rb, o), oq = __shfl_xor_sync(0xffffffffu, rq, o);
			if (oq >= 0 && (rq < 0 || ov > rb || (ov == rb && oq > rq))) { rb = ov; rq = oq; }
		}
	}
	int best_m, best_q, score;
	nw_start_cell(g, cb, ci, ws.lastD(), rb, rq, &best_m, &best_q, &score);
	nw_walk_warp(g, E, best_m, best_q, s, rows, tseq, t_s, q);
	__syncwarp();
	s.score = score; s.pos = 0;
	*out = s;
	return NW_OK;
}
#endif
