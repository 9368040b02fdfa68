// Host model of the register-systolic wavefront schedule of simsandgames_b200/csrc/sor_systolic.cu.
// Test infrastructure (tests/test_systolic_model.py builds and runs it on the CPU): it executes the SAME per-lane dataflow
// as the CUDA kernel -- strips of 32 columns skewed one column per stage, lane = column mod 32, K pipeline stages per lane,
// rotate "shuffles" for the row-direction faces, the wrap lane taking its hand-over from the neighbour strip's ring -- with
// the warp's lanes as array elements and tasks run one after the other, and checks the result bit for bit against the plain
// lexicographic Gauss-Seidel loops of flip_fluid.h:458-494 (i outer, j inner).
//
//   g++ -O2 -ffp-contract=off -o systolic_model systolic_model.cpp && ./systolic_model
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Grid {
	int nf, ns;
	std::vector<float> vf, vs, p, drift;
	std::vector<uint32_t> code;  // bit0 active, bit1..4: s != 0 at fast-, fast+, slow-, slow+
};

static inline void cell_update(bool xfast, uint32_t cd, float dr, float cp, float omega, float& vfl, float& vfh, float& vsl, float& vsh, float& p) {
	if (!(cd & 1u)) return;
	float div = xfast ? ((vfh - vfl) + vsh) - vsl : ((vsh - vsl) + vfh) - vfl;
	div -= dr;
	if (cd == 31u) {
		float pv = (-div * 0.25f) * omega;
		p = p + cp * pv;
		vfl = vfl - pv; vfh = vfh + pv; vsl = vsl - pv; vsh = vsh + pv;
	} else {
		float sfm = (cd & 2u) ? 1.f : 0.f, sfp = (cd & 4u) ? 1.f : 0.f, ssm = (cd & 8u) ? 1.f : 0.f, ssp = (cd & 16u) ? 1.f : 0.f;
		float ssum = xfast ? ((sfm + sfp) + ssm) + ssp : ((ssm + ssp) + sfm) + sfp;
		float pv = -div / ssum;
		pv *= omega;
		p = p + cp * pv;
		vfl = vfl - sfm * pv; vfh = vfh + sfp * pv; vsl = vsl - ssm * pv; vsh = vsh + ssp * pv;
	}
}

// the reference order: F (x) outer, S (y) inner, in place
static void gauss_seidel(Grid& g, int iters, bool xfast, float cp, float omega) {
	for (int it = 0; it < iters; it++)
		for (int F = 1; F < g.nf - 1; F++)
			for (int S = 1; S < g.ns - 1; S++) {
				size_t c = (size_t)F + (size_t)g.nf * S;
				cell_update(xfast, g.code[c], g.drift[c], cp, omega, g.vf[c], g.vf[c + 1], g.vs[c], g.vs[c + g.nf], g.p[c]);
			}
}

// ---- the systolic schedule
constexpr int W = 32;
struct Hand { float p, vs, drift; uint32_t code; };
struct Ring {                       // what strip w leaves for strip w+1, per grid row (+ guard rows)
	int K, rows;
	std::vector<float> vf;          // [K][rows]   vf_hi' of the strip's last column at stage k
	std::vector<Hand> hand;         // [K][rows]   stage k's output of that column (hand-over to stage k+1 of the next strip)
	void init(int K_, int ns) { K = K_; rows = ns + 2 * W + 2 * K_ + 8; vf.assign((size_t)K * rows, 0.f); hand.assign((size_t)K * rows, Hand{0, 0, 0, 0}); }
	int at(int k, int row) const { return k * rows + (row + W + K); }   // rows -W-K .. ns+W+K are addressable
};

template <int K>
static void run_task(Grid& g, int w, bool xfast, float cp, float omega, const Ring* in, Ring* out) {
	const int nf = g.nf, ns = g.ns;
	// per lane, per stage registers
	float vs_lo[K][W] = {}, vf_lo_in[K][W] = {}, vf_hi_out[K][W] = {};
	// A[k]: outputs of stage k-1 from the previous step (row S_k + 1); B[k]: those from two steps ago (row S_k).  k = 1..K-1
	float A_p[K][W] = {}, A_vs[K][W] = {}, A_dr[K][W] = {}, A_vfl[K][W] = {};
	uint32_t A_cd[K][W] = {};
	float B_p[K][W] = {}, B_dr[K][W] = {};
	uint32_t B_cd[K][W] = {};
	auto load = [&](const std::vector<float>& a, int F, int S) { return (F >= 0 && F < nf && S >= 0 && S < ns) ? a[(size_t)F + (size_t)nf * S] : 0.f; };
	auto loadc = [&](int F, int S) { return (F >= 0 && F < nf && S >= 0 && S < ns) ? g.code[(size_t)F + (size_t)nf * S] : 0u; };
	const int T = ns + (W - 1) + (K - 1) + 1;
	for (int t = 0; t < T; t++) {
		float o_p[K][W], o_vsl[K][W], o_vfl[K][W], o_dr[K][W];
		uint32_t o_cd[K][W];
		for (int k = 0; k < K; k++) {
			// "shuffles": every lane reads its neighbours' outputs of the previous step
			float vf_hi_in[W];
			for (int l = 0; l < W; l++) {
				const int c = (l + k) & 31, F = W * w - k + c, S = t - c - k;
				vf_hi_in[l] = k == 0 ? load(g.vf, F + 1, S) : A_vfl[k][(l + 1) & 31];
			}
			for (int l = 0; l < W; l++) {
				const int c = (l + k) & 31, F = W * w - k + c, S = t - c - k;
				float vfl = vf_lo_in[k][l], vfh = vf_hi_in[l], vsl = vs_lo[k][l], vsh, pp, dr;
				uint32_t cd;
				if (k == 0) { vsh = load(g.vs, F, S + 1); pp = load(g.p, F, S); dr = load(g.drift, F, S); cd = loadc(F, S); }
				else { vsh = A_vs[k][l]; pp = B_p[k][l]; dr = B_dr[k][l]; cd = B_cd[k][l]; }
				cell_update(xfast, cd, dr, cp, omega, vfl, vfh, vsl, vsh, pp);
				o_p[k][l] = pp; o_vsl[k][l] = vsl; o_vfl[k][l] = vfl; o_dr[k][l] = dr; o_cd[k][l] = cd;
				vs_lo[k][l] = vsh;          // carried to the next row
				vf_hi_out[k][l] = vfh;
				(void)F; (void)S;
			}
		}
		// end of step: rotate the row-direction faces, shift the hand-over queues, exchange with the rings at the wrap lanes
		for (int k = 0; k < K; k++) {
			float rot[W];
			for (int l = 0; l < W; l++) rot[l] = vf_hi_out[k][(l + 31) & 31];
			for (int l = 0; l < W; l++) {
				const int c = (l + k) & 31, S = t - c - k;          // this lane's row in the step just done
				vf_lo_in[k][l] = rot[l];
				if (c == 31) out->vf[out->at(k, S)] = vf_hi_out[k][l];   // last column of the strip: for the next strip's first column
				if (c == 0) vf_lo_in[k][l] = in ? in->vf[in->at(k, S + 1)] : 0.f;   // first column: its row of the NEXT step
			}
		}
		for (int k = K - 1; k >= 1; k--)
			for (int l = 0; l < W; l++) {
				B_p[k][l] = A_p[k][l]; B_dr[k][l] = A_dr[k][l]; B_cd[k][l] = A_cd[k][l];
				A_p[k][l] = o_p[k - 1][l]; A_vs[k][l] = o_vsl[k - 1][l]; A_dr[k][l] = o_dr[k - 1][l]; A_cd[k][l] = o_cd[k - 1][l];
				A_vfl[k][l] = o_vfl[k - 1][l];
				const int c1 = (l + k - 1) & 31;                      // position of this lane in stage k-1
				if (c1 == 31) {                                       // ... it is the first column of stage k: hand-over crosses the strips
					const int S1 = t - 31 - (k - 1);                  // row stage k-1 just did here
					out->hand[out->at(k - 1, S1)] = Hand{A_p[k][l], A_vs[k][l], A_dr[k][l], A_cd[k][l]};
					const int Sin = t + 2 - k;                        // stage k does row t+1-k here next step and needs the record of the row above it
					Hand h = in ? in->hand[in->at(k - 1, Sin)] : Hand{0, 0, 0, 0};
					A_p[k][l] = h.p; A_vs[k][l] = h.vs; A_dr[k][l] = h.drift; A_cd[k][l] = h.code;
				}
			}
		// final stage: results of the chunk go back to the grid
		for (int l = 0; l < W; l++) {
			const int k = K - 1, c = (l + k) & 31, F = W * w - k + c, S = t - c - k;
			if (F < 1 || F > nf - 1 || S < 1 || S > ns - 1) continue;
			size_t ci = (size_t)F + (size_t)nf * S;
			if (F <= nf - 2 && S <= ns - 2) g.p[ci] = o_p[k][l];
			if (S <= ns - 2) g.vf[ci] = o_vfl[k][l];
			if (F <= nf - 2) g.vs[ci] = o_vsl[k][l];
		}
	}
}

template <int K>
static void systolic(Grid& g, int chunks, bool xfast, float cp, float omega) {
	const int NW = (g.nf + K - 1 + W - 1) / W;
	std::vector<Ring> ring((size_t)NW * 2);
	for (auto& r : ring) r.init(K, g.ns);
	// tasks in key order w + 2q (every dependency has a smaller key)
	for (int key = 0; key <= (NW - 1) + 2 * (chunks - 1); key++)
		for (int q = 0; q < chunks; q++) {
			int w = key - 2 * q;
			if (w < 0 || w >= NW) continue;
			Ring* out = &ring[(size_t)w * 2 + (q & 1)];
			std::fill(out->vf.begin(), out->vf.end(), 0.f);
			std::fill(out->hand.begin(), out->hand.end(), Hand{0, 0, 0, 0});
			const Ring* in = w > 0 ? &ring[(size_t)(w - 1) * 2 + (q & 1)] : nullptr;
			run_task<K>(g, w, xfast, cp, omega, in, out);
		}
}

static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return (float)((rng_state >> 8) & 0xffff) / 65536.f - 0.5f; }

static Grid make_grid(int nf, int ns, int pattern) {
	Grid g;
	g.nf = nf; g.ns = ns;
	size_t n = (size_t)nf * ns;
	g.vf.resize(n); g.vs.resize(n); g.p.assign(n, 0.f); g.drift.assign(n, 0.f); g.code.assign(n, 0);
	std::vector<float> s(n, 1.f);
	std::vector<int> fluid(n, 1);
	for (int S = 0; S < ns; S++)
		for (int F = 0; F < nf; F++) {
			size_t c = (size_t)F + (size_t)nf * S;
			g.vf[c] = frand(); g.vs[c] = frand();
			if (F == 0 || F == nf - 1 || S == 0 || S == ns - 1) s[c] = 0.f;
			if (pattern >= 1) {  // a disc obstacle and an air region
				float dx = F - 0.6f * nf, dy = S - 0.55f * ns;
				if (dx * dx + dy * dy < 0.02f * nf * ns) s[c] = 0.f;
				if (S < ns / 5) fluid[c] = 0;
			}
			if (pattern == 2 && ((F * 7 + S * 13) % 11 == 0)) s[c] = 0.f;   // scattered solids: every code value occurs
			if (pattern == 3 && (F == nf - 1 || S == ns - 1) && ((F + S) & 1)) s[c] = 1.f;   // open border cells: the border faces change
			if (pattern >= 1 && (F + 2 * S) % 5 == 0) g.drift[c] = 0.25f * (frand() + 0.5f);
		}
	for (int S = 1; S < ns - 1; S++)
		for (int F = 1; F < nf - 1; F++) {
			size_t c = (size_t)F + (size_t)nf * S;
			if (s[c] == 0.f || !fluid[c]) continue;
			uint32_t m = (s[c - 1] != 0.f ? 2u : 0u) | (s[c + 1] != 0.f ? 4u : 0u) | (s[c - nf] != 0.f ? 8u : 0u) | (s[c + nf] != 0.f ? 16u : 0u);
			if (m) g.code[c] = m | 1u;
		}
	for (size_t c = 0; c < n; c++) if (!(g.code[c] & 1u)) g.drift[c] = 0.f;
	return g;
}

static bool same(const std::vector<float>& a, const std::vector<float>& b, const char* what, int nf) {
	for (size_t i = 0; i < a.size(); i++)
		if (memcmp(&a[i], &b[i], 4) != 0) {
			printf("  %s differs at F=%d S=%d: %.9g vs %.9g\n", what, (int)(i % nf), (int)(i / nf), a[i], b[i]);
			return false;
		}
	return true;
}

template <int K>
static bool check(int nf, int ns, int chunks, int pattern, bool xfast) {
	Grid a = make_grid(nf, ns, pattern), b = a;
	const float cp = 17.5f, omega = 1.9f;
	gauss_seidel(a, K * chunks, xfast, cp, omega);
	systolic<K>(b, chunks, xfast, cp, omega);
	bool ok = same(a.vf, b.vf, "vel_f", nf) && same(a.vs, b.vs, "vel_s", nf) && same(a.p, b.p, "p", nf);
	printf("%s  K=%d chunks=%d grid %dx%d pattern %d xfast=%d\n", ok ? "ok  " : "FAIL", K, chunks, nf, ns, pattern, (int)xfast);
	return ok;
}

int main() {
	bool ok = true;
	ok &= check<1>(37, 29, 1, 0, true);
	ok &= check<1>(37, 29, 3, 2, true);
	ok &= check<2>(40, 33, 2, 1, false);
	ok &= check<5>(134, 101, 2, 1, true);
	ok &= check<5>(97, 45, 3, 2, false);
	ok &= check<10>(134, 101, 1, 2, true);
	ok &= check<10>(70, 130, 3, 3, true);
	ok &= check<10>(33, 8, 2, 3, false);
	ok &= check<10>(3, 3, 1, 0, true);
	ok &= check<5>(64, 64, 4, 3, true);
	printf(ok ? "ALL OK\n" : "FAILED\n");
	return ok ? 0 : 1;
}
