// TEST INFRASTRUCTURE: lock-step CPU emulation of the warp wavefront in kma_b200/csrc/kmagpu_nw.cuh.
// The per-lane step / start-cell / walk functions are the very source the CUDA kernel compiles; only the
// shuffles are replaced by a snapshot of the neighbour lane's registers. `order` runs the 32 lanes of a step in
// ascending (0) or descending (1) order: a same-step memory hazard shows up as a difference between the two.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../kma_b200/csrc/kmagpu_nw.cuh"

extern "C" int emu_nw(const int *pen29, const uint64_t *tseq, const uint8_t *query, int k, int t_s, int t_e, int q_s,
                      int q_e, int band, int order, int d8, int *out6, long long *steps) {
	NwPen pen;
	pen.W1 = pen29[0]; pen.U = pen29[1]; pen.MM = pen29[2]; pen.M = pen29[3];
	memcpy(pen.d, pen29 + 4, 100);
	pen.d8 = d8;
	const int t_len = t_e - t_s, q_len = q_e - q_s;
	NwStat s;
	if (nw_trivial(pen, t_len, q_len, s)) { memcpy(out6, &s, 24); return 0; }
	NwGeo g;
	if (!nw_geo_init(g, pen, t_len, q_len, k, band)) return 2;
	std::vector<uint8_t> E(g.ebytes(), 0xEE);
	std::vector<NwRow> rowbuf(q_len + NW_RING + 1, NwRow{0x3fffffff, 0x3fffffff});
	std::vector<int> lastD(q_len + 1, 0x3fffffff);
	NwLane L[32];
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
	int cb = g.NEG, ci = 0x7fffffff;
	for (int l = 0; l < 32; ++l)
		if (L[l].colBest > cb || (L[l].colBest == cb && L[l].colBestI < ci)) { cb = L[l].colBest; ci = L[l].colBestI; }
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
