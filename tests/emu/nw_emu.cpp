// TEST INFRASTRUCTURE: lock-step CPU emulation of the warp wavefront in kma_b200/csrc/kmagpu_nw.cuh.
// The per-lane step / start-cell / walk functions are the very source the CUDA kernel compiles; only the
// shuffles are replaced by a snapshot of the neighbour lane's registers. `order` runs the 32 lanes of a step in
// ascending (0) or descending (1) order: a same-step memory hazard shows up as a difference between the two.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../kma_b200/csrc/kmagpu_nw.cuh"

// Row sweep (nw_rs_*): the phases of one row run for all 32 lanes in turn, the shuffles between them are array reads.
template <int C, bool BANDED>
static int emu_rs_fill(const NwGeo &g, const NwPen &pen, const uint64_t *tseq, int t_s, const uint8_t *q, uint8_t *E, int *lastD,
                       int *cbo, int *cio) {
	NwRsLane<C> L[32];
	unsigned long long tab[5];
	for (int tn = 0; tn < 5; ++tn) tab[tn] = nw_rs_tab(pen, tn);
	const uint8_t *qlast = q + g.q_len - 1;
	for (int l = 0; l < 32; ++l) nw_rs_init<C, BANDED>(g, L[l], l, qlast);
	const int Ue = g.W1 > g.U ? g.W1 : g.U, step = C * Ue;
	int bad = 0;
	int i1 = 0, i2 = 0;
	if (BANDED) nw_rs_interior(g, &i1, &i2);
	const int Qs_in = nw_max(g.NEG + g.W1, g.NEG + g.U);
	for (int i = 0; i < g.t_len; ++i) {
		int nbD[32], nbP[32], B[32], old[32], Qf[32], e[32][C], Dlast[32], Qlast[32], fP[32][C];
		if (BANDED && i >= i1 && i < i2) {   // the interior rows of a band: the lean passes of the kernel
			const unsigned long long tr = tab[nw_nuc(tseq, t_s + g.t_len - 1 - i)];
			for (int l = 0; l < 32; ++l) { const int s = l < 31 ? l + 1 : 31; nbD[l] = L[s].pD[0]; nbP[l] = L[s].pP[0]; }
			for (int l = 0; l < 32; ++l) {
				const int ubc = (g.band >= l * C && g.band < l * C + C) ? g.band - l * C : -1;
				B[l] = nw_rs_pass1i<C>(g, L[l], l == 0, ubc, (unsigned)tr, (unsigned)(tr >> 32), nbD[l], nbP[l], Ue, Qs_in, fP[l]) - l * step;
			}
			for (int o = 1; o < 32; o <<= 1) {
				memcpy(old, B, sizeof(B));
				for (int l = o; l < 32; ++l) if (old[l - o] > B[l]) B[l] = old[l - o];
			}
			for (int l = 0; l < 32; ++l) Qf[l] = l ? B[l - 1] + (l - 1) * step : NW_NINF;
			for (int l = 0; l < 32; ++l) {
				const int ubc = (g.band >= l * C && g.band < l * C + C) ? g.band - l * C : -1;
				nw_rs_pass2i<C>(g, L[l], l == 0, ubc, Qf[l], Qs_in, fP[l], e[l], &Dlast[l], &Qlast[l]);
			}
			for (int l = 0; l < 32; ++l) { const int s = l ? l - 1 : 0; e[l][0] = nw_rs_fix0i(g, l == 0, e[l][0], Dlast[s], Qlast[s]); }
			for (int l = 0; l < 32; ++l)
				for (int c = 0; c < C; ++c) E[(size_t)i * (32 * C) + l * C + c] = (uint8_t)e[l][c];
			for (int l = 0; l < 32; ++l) {
				const int q8n = nw_rs_q8<BANDED>(g, i + 1, l * C + C - 1, qlast);
				for (int c = 0; c + 1 < C; ++c) L[l].qs[c] = L[l].qs[c + 1];
				L[l].qs[C - 1] = q8n;
			}
			continue;
		}
		NwRsRow R;
		nw_rs_row(g, R, i, tab[nw_nuc(tseq, t_s + g.t_len - 1 - i)]);
		for (int l = 0; l < 32; ++l) {
			if (BANDED) { const int s = l < 31 ? l + 1 : 31; nbD[l] = L[s].pD[0]; nbP[l] = L[s].pP[0]; }
			else { const int s = l ? l - 1 : 0; nbD[l] = L[s].pD[C - 1]; nbP[l] = 0; }
		}
		for (int l = 0; l < 32; ++l) B[l] = nw_rs_pass1<C, BANDED>(g, L[l], l, R, nbD[l], nbP[l], Ue, fP[l]) - l * step;
		for (int o = 1; o < 32; o <<= 1) {
			memcpy(old, B, sizeof(B));
			for (int l = o; l < 32; ++l) if (old[l - o] > B[l]) B[l] = old[l - o];
		}
		for (int l = 0; l < 32; ++l) Qf[l] = l ? B[l - 1] + (l - 1) * step : NW_NINF;
		for (int l = 0; l < 32; ++l) nw_rs_pass2<C, BANDED>(g, L[l], l, R, Qf[l], fP[l], e[l], &Dlast[l], &Qlast[l]);
		for (int l = 0; l < 32; ++l) {
			const int s = l ? l - 1 : 0;
			const int u = l * C;
			if (u > R.ua && u <= R.ub) {   // the scan must reproduce the sequential recurrence
				const int a = Dlast[s] + g.W1, b = Qlast[s] + g.U;
				if (Qf[l] != (a < b ? b : a)) ++bad;
			}
			e[l][0] = nw_rs_fix0<C>(g, l, R, e[l][0], Dlast[s], Qlast[s]);
		}
		for (int l = 0; l < 32; ++l)
			for (int c = 0; c < C; ++c) E[(size_t)i * (32 * C) + l * C + c] = (uint8_t)e[l][c];
		for (int l = 0; l < 32; ++l) nw_rs_rowend<C, BANDED>(g, L[l], l, R, i, nw_rs_q8<BANDED>(g, i + 1, l * C + C - 1, qlast));
	}
	for (int l = 0; l < 32; ++l) nw_rs_lastrow<C, BANDED>(g, L[l], l, lastD);
	int cb = g.NEG, ci = 0x7fffffff;
	for (int l = 0; l < 32; ++l)
		if (L[l].colBest > cb || (L[l].colBest == cb && L[l].colBestI < ci)) { cb = L[l].colBest; ci = L[l].colBestI; }
	*cbo = cb; *cio = ci;
	return bad;
}

template <bool BANDED>
static int emu_rs_dispatch(const NwGeo &g, const NwPen &pen, const uint64_t *tseq, int t_s, const uint8_t *q, uint8_t *E, int *lastD,
                           int *cb, int *ci) {
	switch (g.C) {
	case 1: return emu_rs_fill<1, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	case 2: return emu_rs_fill<2, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	case 3: return emu_rs_fill<3, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	case 4: return emu_rs_fill<4, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	case 6: return emu_rs_fill<6, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	case 8: return emu_rs_fill<8, BANDED>(g, pen, tseq, t_s, q, E, lastD, cb, ci);
	}
	return -1;
}

// rs = 1: rows of up to 256 cells run as the row sweep (what the kernel does), rs = 0: the wavefront for every size.
// emap (optional, t_len * q_len bytes): the traceback byte of every cell inside the matrix / band, 0xFF outside.
extern "C" int emu_nw2(const int *pen29, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s,
                       int q_e, int band, int order, int d8, int rs, int *out6, long long *steps, uint8_t *emap) {
	NwPen pen;
	pen.W1 = pen29[0]; pen.U = pen29[1]; pen.MM = pen29[2]; pen.M = pen29[3];
	memcpy(pen.d, pen29 + 4, 100);
	pen.d8 = d8;
	const int t_len = t_e - t_s, q_len = q_e - q_s;
	NwStat s;
	if (nw_trivial(pen, t_len, q_len, s)) { memcpy(out6, &s, 24); return 0; }
	NwGeo g;
	if (!nw_geo_init(g, pen, t_len, q_len, k, band, rs ? NW_RS_MAXC : 0)) return 2;
	std::vector<uint8_t> E(g.ebytes(), 0xEE);
	std::vector<NwRow> rowbuf(q_len + NW_RING + 1, NwRow{0x3fffffff, 0x3fffffff});
	std::vector<int> lastD(q_len + 1, 0x3fffffff);
	NwLane L[32];
	int cb = g.NEG, ci = 0x7fffffff;
	if (g.C) {
		const int bad = g.banded ? emu_rs_dispatch<true>(g, pen, tseq, t_s, query + q_s, E.data(), lastD.data(), &cb, &ci)
		                         : emu_rs_dispatch<false>(g, pen, tseq, t_s, query + q_s, E.data(), lastD.data(), &cb, &ci);
		if (bad) return 3;
		if (steps) *steps += (long long)g.t_len * g.C;
	} else {
	for (int l = 0; l < 32; ++l) nw_lane_init(g, pen, L[l], l, tseq, t_s, query + q_s);
	for (int T = 0; T < g.Tmax; ++T) {
		int aD[32], aP[32];
		for (int l = 0; l < 32; ++l) { aD[l] = L[l ? l - 1 : 0].myD; aP[l] = L[l ? l - 1 : 0].myP; }
		for (int x = 0; x < 32; ++x) {
			const int l = order ? 31 - x : x;
			nw_lane_step(g, pen, L[l], l, aD[l], aP[l], tseq, t_s, query + q_s + q_len - 1, E.data() + (size_t)T * 32, rowbuf.data(), lastD.data());
		}
		nw_lane0_prefetch(g, L[0], rowbuf.data());
	}
	if (steps) *steps += g.Tmax;
	for (int l = 0; l < 32; ++l)
		if (L[l].colBest > cb || (L[l].colBest == cb && L[l].colBestI < ci)) { cb = L[l].colBest; ci = L[l].colBestI; }
	}
	if (emap)
		for (int i = 0; i < t_len; ++i)
			for (int j = 0; j < q_len; ++j)
				emap[(size_t)i * q_len + j] = (j >= g.jlo(i) && j <= g.jhi(i)) ? E[g.eaddr(i, j)] : 0xFF;
	int rb = g.NEG, rq = -1;
	if (k == -2) {
		int qlo, qhi;
		nw_row0_range(g, &qlo, &qhi);
		for (int qp = qlo; qp <= qhi; ++qp) { int v = lastD[q_len - 1 - qp]; if (rq < 0 || v >= rb) { rb = v; rq = qp; } }
	}
	int bm, bq, sc;
	nw_start_cell(g, cb, ci, lastD.data(), rb, rq, &bm, &bq, &sc);
	nw_walk(g, E.data(), bm, bq, s);
	s.score = sc; s.pos = 0;
	memcpy(out6, &s, 24);
	return 0;
}

extern "C" int emu_nw(const int *pen29, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s,
                      int q_e, int band, int order, int d8, int *out6, long long *steps) {
	return emu_nw2(pen29, tseq, query, k, t_s, t_e, q_s, q_e, band, order, d8, 1, out6, steps, nullptr);
}

// nw_thread (one thread per problem): estride / rstride > 1 emulate the interleaved device layouts
extern "C" int emu_nw_thread(const int *pen29, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s,
                             int q_e, int d8, int stride, int *out6, uint8_t *emap) {
	NwPen pen;
	pen.W1 = pen29[0]; pen.U = pen29[1]; pen.MM = pen29[2]; pen.M = pen29[3];
	memcpy(pen.d, pen29 + 4, 100);
	pen.d8 = d8;
	const int t_len = t_e - t_s, q_len = q_e - q_s;
	NwStat s;
	if (nw_trivial(pen, t_len, q_len, s)) { memcpy(out6, &s, 24); return 0; }
	std::vector<NwRow> rows((size_t)q_len * stride + 1, NwRow{0x3fffffff, 0x3fffffff});
	std::vector<uint8_t> E((size_t)t_len * q_len * stride + 1, 0xEE), qs((size_t)q_len * stride + 1, 0xEE);
	for (int j = 0; j < q_len; ++j) qs[(size_t)j * stride] = (uint8_t)(query[q_s + q_len - 1 - j] << 3);   // what the kernel stages
	unsigned long long tab[5];
	for (int tn = 0; tn < 5; ++tn) tab[tn] = nw_pack_row(pen, tn);
	if (d8) nw_thread<true>(pen, tab, tseq, t_s, t_len, qs.data(), stride, q_len, k, rows.data(), stride, E.data(), (size_t)stride, &s);
	else nw_thread<false>(pen, tab, tseq, t_s, t_len, qs.data(), stride, q_len, k, rows.data(), stride, E.data(), (size_t)stride, &s);
	if (emap)
		for (int c = 0; c < t_len * q_len; ++c) emap[c] = E[(size_t)c * stride];
	memcpy(out6, &s, 24);
	return 0;
}
