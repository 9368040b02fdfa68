// Blocked-wavefront Gauss-Seidel SOR in shared memory (FLIP_SOLVER_LEX_WAVEFRONT, the default): the reference's lexicographic sweep order
// (flip_fluid.h:458-494 / fluid.h:106-138), bit for bit, staged through shared memory with
// temporal blocking, as ONE persistent kernel per solve.
//
// Why this order and not plain red-black: with the reference's iteration count the lexicographic sweep
// carries pressure information across the whole grid in the sweep direction every iteration; red-black
// moves it two cells per iteration and, at the reference's 50 iterations, leaves tall liquid columns
// unsupported (the dam-break configs blow up within ~20 steps, profiles/r1_stability.md).  The
// anti-diagonal wavefront is the parallel schedule that IS the lexicographic order: a cell shares faces
// only with its four edge neighbours, so all cells with i+j = const are independent and iteration k+1 may
// trail iteration k by two diagonals.
//
// Blocking.  Let (F,S) be interior cell coordinates along the fast/slow storage axes.  The iteration
// space (F,S,k) is cut into tiles  { (F0-k+a, S0-k+b, q*Kc+k) : 0<=a,b<T, 0<=k<Kc }  -- squares of side T
// skewed by one cell per iteration.  In the skewed coordinates every dependency of (a,b,k) points to
// (a-1,b,.), (a,b-1,.) or (a-1,b-1,k-1), so
//   * inside a tile, ALL Kc iterations of local diagonal a+b = t are independent: 2T-1 block-wide steps
//     do T*T*Kc cell updates out of shared memory;
//   * between tiles the dependencies are (I-1,J,q), (I,J-1,q) and (I+1,J+1,q-1): tiles are handed out in
//     wave order I+J+3q from an atomic counter and wait on completion flags of earlier tiles only.
// HBM traffic per solve is ~(1+2Kc/T) grid sweeps per chunk of Kc iterations instead of Kc sweeps, with
// no redundant computation (skewing, not halo recompute).  Compiled with -fmad=false.
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace fg {

#ifndef SOR_TILE
#define SOR_TILE 32
#endif
constexpr int kTile = SOR_TILE;  // T: side of the skewed square (32: one warp spans a tile row).  -DSOR_TILE=64 (1024 threads, one CTA per SM,
                                 // half the tile waves, twice the diagonal steps per tile) measured 5.18 vs 4.68 ms at S2 and 40.6 vs 40.2 ms at S4
constexpr int kMaxChunk = 16;    // Kc <= kMaxChunk <= T
constexpr int kBoxW = kTile + kMaxChunk + 2;  // 50: lanes of one iteration stride by W-1 = 49 (odd, conflict-free) and lanes of
                                             // different iterations packed into one warp collide only for (da,dk) = (2,10), (4,-12) ...

struct TiledArgs {
	int nf, ns;
	float* vel_f;
	float* vel_s;
	float* p;
	const unsigned char* code;  // bit0 active, bit1..4 = s != 0 at fast-, fast+, slow-, slow+ neighbour
	const float* drift;         // max(0, density-rest) when drift compensation applies, else nullptr
	const int4* tiles;          // (I, J, q, kq) in wave order
	const unsigned char* act;   // act[J0*na + I0] != 0: the unskewed TxT square (I0,J0) holds a cell the solver updates
	int na, nb;
	int n_tiles, ni, nj, kc;
	int* flags;                 // flags[(q*nj+J)*ni+I] == epoch  <=> tile complete
	int* counter;
	int epoch;
	float cp, omega;
	unsigned long long* prof;  // [0] wait [1] load [2] compute [3] write-back cycles (thread 0 of every CTA), [4] tiles
	int* status;               // device error word: a dependency wait that timed out stores FLIPGPU_ESTATE here ...
	volatile int* host_status; // ... and in this word of pinned host memory (written by the kernel: no D2H copy on the stream)
};

// ceil(2^16 / d): floor(n / d) == (n * magic) >> 16 for every n < 1024, d <= 64 (checked exhaustively)
__device__ __forceinline__ unsigned div_magic(int d) { return (65536u + (unsigned)d - 1u) / (unsigned)d; }

// The per-step item decode is a multiply-shift by a constant-bank magic number, every operand of a cell is fetched in ONE
// shared-memory round trip, and the per-solve planes (cell code, drift term) plus the activity map are fetched before the
// predecessor flags are awaited (r2 A/B on B200, profiles/r2_ab_v2_variants_4096.jsonl: 5.19 -> 4.46 ms at S2, bit-identical).
template <bool XFAST, bool DRIFT, int KW>  // KW warps per CTA (>= iterations per chunk)
__global__ void __launch_bounds__(kTile * KW, (kTile * KW > 512 ? 1 : (KW <= 8 ? 4 : 2))) sor_gs_tiled_kernel(TiledArgs A) {
	// dynamic shared memory (> 48 KB with the drift term): four float planes and the cell code as a 32-bit word per cell
	// (a byte plane would put several lanes of a diagonal into one bank)
	extern __shared__ float smem_tile[];
	float* s_vf = smem_tile;
	float* s_vs = s_vf + kBoxW * kBoxW;
	float* s_p = s_vs + kBoxW * kBoxW;
	unsigned* s_code = reinterpret_cast<unsigned*>(s_p + kBoxW * kBoxW);
	float* s_dr = reinterpret_cast<float*>(s_code + kBoxW * kBoxW);  // only touched when DRIFT
	__shared__ int s_tile;
	__shared__ int s_ok;
	__shared__ unsigned s_magic[kTile + 1];
	const int tid = threadIdx.y * kTile + threadIdx.x;
	if (tid >= 1 && tid <= kTile) s_magic[tid] = div_magic(tid);
	const int nthreads = kTile * blockDim.y;
	const int nf = A.nf, ns = A.ns;
	for (;;) {
		__syncthreads();
		if (tid == 0) s_tile = atomicAdd(A.counter, 1);
		__syncthreads();
		const int ti = s_tile;
		if (ti >= A.n_tiles) return;
		const int4 tl = A.tiles[ti];
		const int I = tl.x, J = tl.y, q = tl.z, kq = tl.w;
		long long c0 = clock64();
		// the activity of the tile's squares does not depend on the predecessors -- fetch it while waiting for them.  The skewed
		// tile lies inside the four unskewed squares (I-1..I, J-1..J): if none of them holds an active cell (all air / solid)
		// the tile changes nothing and is published as soon as its predecessors are done (dam-break: about half the tiles)
		bool any = false;
		for (int jj = max(J - 1, 0); jj <= min(J, A.nb - 1); jj++)
			for (int ii = max(I - 1, 0); ii <= min(I, A.na - 1); ii++) any |= A.act[jj * A.na + ii] != 0;
		const int F0 = 1 + I * kTile, S0 = 1 + J * kTile;
		const int FB = F0 - A.kc + 1, SB = S0 - A.kc + 1;  // box origin (cells FB..FB+kBoxW-1)
		if (any) {  // ... and so are the per-solve cell constants: stage the code / drift planes of the box now
			constexpr int kRoundsP = (kBoxW * kBoxW + kTile * KW - 1) / (kTile * KW);
			unsigned p_cd[kRoundsP];
			float p_dr[kRoundsP];
#pragma unroll
			for (int r = 0; r < kRoundsP; r++) {
				int e = tid + r * nthreads;
				int F = FB + e % kBoxW, S = SB + e / kBoxW;
				p_cd[r] = 0; p_dr[r] = 0.f;
				if (e < kBoxW * kBoxW && F >= 0 && F < nf && S >= 0 && S < ns) {
					size_t c = (size_t)F + (size_t)nf * S;
					p_cd[r] = A.code[c];
					if (DRIFT) p_dr[r] = A.drift[c];
				}
			}
#pragma unroll
			for (int r = 0; r < kRoundsP; r++) {
				int e = tid + r * nthreads;
				if (e < kBoxW * kBoxW) {
					s_code[e] = p_cd[r];
					if (DRIFT) s_dr[e] = p_dr[r];
				}
			}
		}
		if (tid == 0) {
			// the three predecessor flags are polled together (one L2 round trip per poll instead of up to three in a row),
			// without sleeping: one thread per CTA spins, and the tile usually sits on the critical path of the wavefront.
			// The spin is bounded (~seconds); a time-out raises the solver's error word and ends the launch -- the host
			// reports FLIPGPU_ESTATE at its next synchronising call instead of returning a stale pressure field.
			const int* f1 = I > 0 ? A.flags + (q * A.nj + J) * A.ni + I - 1 : nullptr;
			const int* f2 = J > 0 ? A.flags + (q * A.nj + J - 1) * A.ni + I : nullptr;
			const int* f3 = q > 0 ? A.flags + ((q - 1) * A.nj + min(J + 1, A.nj - 1)) * A.ni + min(I + 1, A.ni - 1) : nullptr;
			int ok = 0;
			for (long long spins = 0; spins < (1ll << 27); spins++) {
				int v1 = f1 ? *((volatile const int*)f1) : A.epoch, v2 = f2 ? *((volatile const int*)f2) : A.epoch,
				    v3 = f3 ? *((volatile const int*)f3) : A.epoch;
				if (v1 == A.epoch && v2 == A.epoch && v3 == A.epoch) { ok = 1; break; }
				if ((spins & 1023) == 1023 && *((volatile const int*)A.status) != 0) break;  // another CTA already gave up
			}
			if (!ok) { atomicExch(A.status, FLIPGPU_ESTATE); *A.host_status = FLIPGPU_ESTATE; }
			s_ok = ok;
			__threadfence();
		}
		__syncthreads();
		if (!s_ok) return;
		long long c1 = clock64();
		if (!any) {
			if (tid == 0) *((volatile int*)(A.flags + (q * A.nj + J) * A.ni + I)) = A.epoch;
			continue;
		}
		// ---- stage the box (L2-coherent loads: the producers are other CTAs of this launch).  All loads of a thread are
		// issued before any is consumed, so the box costs one L2 round trip instead of one per element: the tile sits on
		// the critical path of the wavefront and its fixed costs are paid once per wave.
		{
			constexpr int kRounds = (kBoxW * kBoxW + kTile * KW - 1) / (kTile * KW);
			float r_vf[kRounds], r_vs[kRounds], r_p[kRounds];
#pragma unroll
			for (int r = 0; r < kRounds; r++) {
				int e = tid + r * nthreads;
				int lf = e % kBoxW, ls = e / kBoxW;
				int F = FB + lf, S = SB + ls;
				r_vf[r] = 0.f; r_vs[r] = 0.f; r_p[r] = 0.f;
				if (e < kBoxW * kBoxW && F >= 0 && F < nf && S >= 0 && S < ns) {
					size_t c = (size_t)F + (size_t)nf * S;
					r_vf[r] = __ldcg(A.vel_f + c); r_vs[r] = __ldcg(A.vel_s + c); r_p[r] = __ldcg(A.p + c);
				}
			}
#pragma unroll
			for (int r = 0; r < kRounds; r++) {
				int e = tid + r * nthreads;
				if (e < kBoxW * kBoxW) { s_vf[e] = r_vf[r]; s_vs[e] = r_vs[r]; s_p[e] = r_p[r]; }
			}
		}
		__syncthreads();
		long long c2 = clock64();
		// ---- 2T-1 local diagonals; on each, every iteration of the chunk in parallel.
		// Lane a owns local column a: it is busy on steps t in [a, a+T-1] and walks down its box column by one
		// row per step.  Cells outside the grid interior carry code 0 (sor_prep_kernel / the box fill), so no
		// bounds tests are needed here.
		// Work items of step t are the cells (a, t-a) of every iteration k < kq with a in [alo, ahi].  They are dealt out
		// densely over the CTA (item = thread id), so on the ramps of the skewed square -- where a diagonal holds only a few
		// cells -- whole warps have nothing to issue instead of every warp issuing for a few lanes.
		for (int t = 0; t < 2 * kTile - 1; t++) {
			const int alo = max(0, t - (kTile - 1)), cnt = min(kTile - 1, t) - alo + 1;
			if (tid < cnt * kq) {
				const int k = (int)(((unsigned)tid * s_magic[cnt]) >> 16);  // tid / cnt
				const int a = alo + (tid - k * cnt);
				// F = F0-k+a, S = S0-k+(t-a)  ->  box element (F-FB) + W*(S-SB)
				const int e = (a - k + A.kc - 1) + kBoxW * (t - a - k + A.kc - 1);
				// every operand is fetched before the activity test (volatile reads cannot be sunk below it): the seven
				// loads go out back to back, one shared-memory round trip per step instead of two
				const unsigned cd = s_code[e];
				const float vfc = *(volatile float*)(s_vf + e), vfp = *(volatile float*)(s_vf + e + 1), vsc = *(volatile float*)(s_vs + e),
				            vsp = *(volatile float*)(s_vs + e + kBoxW), pp = *(volatile float*)(s_p + e);
				float dr = 0.f;
				if (DRIFT) dr = *(volatile float*)(s_dr + e);
				if (cd & 1u) {
					float div = XFAST ? ((vfp - vfc) + vsp) - vsc : ((vsp - vsc) + vfp) - vfc;
					if (DRIFT) div -= dr;        // the stored term is max(0, density-rest): subtracting 0 is exact
					if (cd == 31u) {             // all four neighbours open: s_sum = 4 and x/4 == x*0.25 exactly
						float pv = (-div * 0.25f) * A.omega;
						s_p[e] = pp + A.cp * pv;
						s_vf[e] = vfc - pv; s_vf[e + 1] = vfp + pv; s_vs[e] = vsc - pv; s_vs[e + kBoxW] = vsp + pv;
					} else {
						float sfm = (cd & 2u) ? 1.f : 0.f, sfp = (cd & 4u) ? 1.f : 0.f;
						float ssm = (cd & 8u) ? 1.f : 0.f, ssp = (cd & 16u) ? 1.f : 0.f;
						float ssum = XFAST ? ((sfm + sfp) + ssm) + ssp : ((ssm + ssp) + sfm) + sfp;
						float pv = -div / ssum;
						pv *= A.omega;
						s_p[e] = pp + A.cp * pv;
						s_vf[e] = vfc - sfm * pv; s_vf[e + 1] = vfp + sfp * pv;
						s_vs[e] = vsc - ssm * pv; s_vs[e + kBoxW] = vsp + ssp * pv;
					}
				}
			}
			__syncthreads();
		}
		long long c3 = clock64();
		// ---- write back exactly what this tile owns: cells it updated, plus their high-side faces
		for (int e = tid; e < kBoxW * kBoxW; e += nthreads) {
			int lf = e % kBoxW, ls = e / kBoxW;
			int F = FB + lf, S = SB + ls;
			if (F < 1 || F > nf - 1 || S < 1 || S > ns - 1) continue;
			int x = F - F0, y = S - S0;
			// cell (x,y) is visited at iterations k in [max(0,-x,-y), min(kq-1, T-1-x, T-1-y)]
			auto visited = [&](int xx, int yy) {
				int lo = max(0, max(-xx, -yy)), hi = min(kq - 1, min(kTile - 1 - xx, kTile - 1 - yy));
				return lo <= hi;
			};
			bool own = visited(x, y) && F <= nf - 2 && S <= ns - 2;
			bool from_f = visited(x - 1, y) && F - 1 >= 1 && S <= ns - 2;  // high-side face of the fast- neighbour
			bool from_s = visited(x, y - 1) && S - 1 >= 1 && F <= nf - 2;  // high-side face of the slow- neighbour
			size_t c = (size_t)F + (size_t)nf * S;
			if (own) __stcg(A.p + c, s_p[e]);
			if (own || from_f) __stcg(A.vel_f + c, s_vf[e]);
			if (own || from_s) __stcg(A.vel_s + c, s_vs[e]);
		}
		// release: every thread has stored, bar.sync, then ONE thread fences (cumulative at gpu scope) and raises the flag
		__syncthreads();
		if (tid == 0) {
			__threadfence();
			*((volatile int*)(A.flags + (q * A.nj + J) * A.ni + I)) = A.epoch;
			if (A.prof) {
				long long c4 = clock64();
				atomicAdd(A.prof + 0, (unsigned long long)(c1 - c0)); atomicAdd(A.prof + 1, (unsigned long long)(c2 - c1));
				atomicAdd(A.prof + 2, (unsigned long long)(c3 - c2)); atomicAdd(A.prof + 3, (unsigned long long)(c4 - c3));
				atomicAdd(A.prof + 4, 1ull);
			}
		}
	}
}

// per-solve constants of every cell: the neighbour-openness code and the drift term
__global__ void __launch_bounds__(256) sor_prep_kernel(SorGrid g, bool drift, unsigned char* __restrict__ code,
                                                       float* __restrict__ drift_out, unsigned char* __restrict__ act, int na,
                                                       size_t c_begin, size_t c_end) {
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	for (size_t c = c_begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < c_end; c += (size_t)gridDim.x * blockDim.x) {
		int a = (int)(c % g.nf), b = (int)(c / g.nf);
		unsigned cd = 0;
		float dr = 0.f;
		if (a >= 1 && a <= g.nf - 2 && b >= 1 && b <= g.ns - 2 && g.cell_type[c] == FLUID_CELL) {
			unsigned m = (g.s[c - 1] != 0.f ? 2u : 0u) | (g.s[c + 1] != 0.f ? 4u : 0u) | (g.s[c - g.nf] != 0.f ? 8u : 0u) |
			             (g.s[c + g.nf] != 0.f ? 16u : 0u);
			if (m) cd = m | 1u;
			if (drift && rest > 0.f) {
				float compression = g.density[c] - rest;
				if (compression > 0.f) dr = compression;
			}
		}
		code[c] = (unsigned char)cd;
		if (drift_out) drift_out[c] = dr;
		if (cd && act) act[((b - 1) / kTile) * na + (a - 1) / kTile] = 1;  // racing writers store the same value
	}
}

struct SorWorkspace {
	int nf = 0, ns = 0, iters = 0, kc = 0, ni = 0, nj = 0, nq = 0, n_tiles = 0, epoch = 0;
	unsigned char* code = nullptr;
	float* drift = nullptr;
	int4* tiles = nullptr;
	int* flags = nullptr;
	int* counter = nullptr;
	unsigned char* act = nullptr;  // activity of the unskewed TxT squares
	int na = 0, nb = 0;
	unsigned long long* prof = nullptr;
	SysWorkspace* sys = nullptr;  // workspace of the register-systolic kernel (sor_systolic.cu, FLIP_SOLVER_LEX_SYSTOLIC)
	int sms = 0;
	int* status = nullptr;    // device error word raised by a dependency wait that timed out ...
	int* h_status = nullptr;  // ... and this word of pinned host memory, written by the kernel itself
};

// diagnostic: cycles spent per tile phase (enabled by FLIPGPU_SOR_PROFILE=1 in the environment at plan time)
int sor_workspace_profile(SorWorkspace* w, unsigned long long out[8]) {
	for (int i = 0; i < 8; i++) out[i] = 0;
	if (w && w->sys && sys_workspace_profile(w->sys, out) == FLIPGPU_OK && out[4] != 0) return FLIPGPU_OK;   // the systolic kernel ran
	if (!w || !w->prof) return FLIPGPU_ESTATE;
	FG_CUDA(cudaMemcpy(out, w->prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	out[5] = (unsigned long long)w->n_tiles; out[6] = (unsigned long long)w->kc; out[7] = (unsigned long long)w->nq;
	return FLIPGPU_OK;
}

void sor_workspace_free(SorWorkspace* w) {
	if (!w) return;
	cudaFree(w->code); cudaFree(w->drift); cudaFree(w->tiles); cudaFree(w->flags); cudaFree(w->counter); cudaFree(w->prof); cudaFree(w->act);
	cudaFree(w->status);
	if (w->h_status) cudaFreeHost(w->h_status);
	sys_workspace_free(w->sys);
	delete w;
}
SysWorkspace** sor_workspace_sys(SorWorkspace** wsp) {
	if (!*wsp) *wsp = new SorWorkspace;
	return &(*wsp)->sys;
}

// Call after the stream that ran the solves has been synchronised: FLIPGPU_ESTATE if an inter-CTA dependency wait of a
// persistent solver kernel timed out (the pressure field of that solve is then incomplete).
int sor_workspace_check(SorWorkspace* w) {
	if (w && w->sys) FG_TRY(sys_workspace_check(w->sys));
	if (!w || !w->h_status || *w->h_status == 0) return FLIPGPU_OK;
	fg::set_error("pressure solve: an inter-CTA dependency wait timed out (GPU preempted or oversubscribed?); the solve is incomplete");
	return FLIPGPU_ESTATE;
}

static int ensure_cell_constants(SorWorkspace* w, const SorGrid& g) {
	if (w->nf != g.nf || w->ns != g.ns) {
		cudaFree(w->code); cudaFree(w->drift);
		w->code = nullptr; w->drift = nullptr;
		size_t n = (size_t)g.nf * g.ns;
		FG_CUDA(cudaMalloc(&w->code, n));
		FG_CUDA(cudaMalloc(&w->drift, n * 4));
		cudaFree(w->act);
		w->na = (g.nf - 2 + kTile - 1) / kTile; w->nb = (g.ns - 2 + kTile - 1) / kTile;
		FG_CUDA(cudaMalloc(&w->act, (size_t)w->na * w->nb));
		w->nf = g.nf; w->ns = g.ns; w->iters = 0;
		if (!w->counter) FG_CUDA(cudaMalloc(&w->counter, 4));
		if (!w->status) {
			FG_CUDA(cudaMalloc(&w->status, 4));
			FG_CUDA(cudaMemset(w->status, 0, 4));
			FG_CUDA(cudaMallocHost(&w->h_status, 4));
			*w->h_status = 0;
		}
		if (!w->prof && getenv("FLIPGPU_SOR_PROFILE")) {
			FG_CUDA(cudaMalloc(&w->prof, 8 * sizeof(unsigned long long)));
			FG_CUDA(cudaMemset(w->prof, 0, 8 * sizeof(unsigned long long)));
		}
		int dev = 0;
		FG_CUDA(cudaGetDevice(&dev));
		FG_CUDA(cudaDeviceGetAttribute(&w->sms, cudaDevAttrMultiProcessorCount, dev));
	}
	return FLIPGPU_OK;
}

static int plan(SorWorkspace* w, const SorGrid& g, int iters, cudaStream_t st) {
	FG_TRY(ensure_cell_constants(w, g));
	if (w->iters != iters) {
		int maxc = kMaxChunk;
		if (const char* e = getenv("FLIPGPU_SOR_KC")) maxc = std::max(1, std::min(kMaxChunk, atoi(e)));  // tuning knob
		int nq = (iters + maxc - 1) / maxc;
		int kc = (iters + nq - 1) / nq;
		int ni = (g.nf - 2 + kc - 1 + kTile - 1) / kTile, nj = (g.ns - 2 + kc - 1 + kTile - 1) / kTile;
		std::vector<int4> tiles;
		tiles.reserve((size_t)ni * nj * nq);
		int wmax = (ni - 1) + (nj - 1) + 3 * (nq - 1);
		for (int wv = 0; wv <= wmax; wv++)
			for (int q = 0; q < nq; q++) {
				int d = wv - 3 * q;
				if (d < 0) break;
				int kq = std::min(kc, iters - q * kc);
				for (int I = std::max(0, d - (nj - 1)); I <= std::min(ni - 1, d); I++) tiles.push_back(make_int4(I, d - I, q, kq));
			}
		cudaFree(w->tiles); cudaFree(w->flags);
		w->tiles = nullptr; w->flags = nullptr;
		FG_CUDA(cudaMalloc(&w->tiles, tiles.size() * sizeof(int4)));
		FG_CUDA(cudaMalloc(&w->flags, (size_t)ni * nj * nq * 4));
		FG_CUDA(cudaMemcpyAsync(w->tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
		FG_CUDA(cudaMemsetAsync(w->flags, 0, (size_t)ni * nj * nq * 4, st));
		FG_CUDA(cudaStreamSynchronize(st));  // tiles is a host temporary
		w->iters = iters; w->kc = kc; w->ni = ni; w->nj = nj; w->nq = nq; w->n_tiles = (int)tiles.size();
		w->epoch = 0;
	}
	return FLIPGPU_OK;
}

// s must be 0/1-valued (checked by the caller); arbitrary float s takes the plain wavefront kernel of sor.cu
int sor_solve_tiled(SorWorkspace** wsp, const SorGrid& g, int iters, float cp, float omega, bool drift, cudaStream_t st,
                    LaunchCounter* lc) {
	if (iters <= 0 || g.nf < 3 || g.ns < 3) return FLIPGPU_OK;
	if (!*wsp) *wsp = new SorWorkspace;
	SorWorkspace* w = *wsp;
	FG_TRY(plan(w, g, iters, st));
	bool use_drift = drift && g.density && g.rest;
	FG_CUDA(cudaMemsetAsync(w->act, 0, (size_t)w->na * w->nb, st));
	sor_prep_kernel<<<wave_grid((size_t)g.nf * g.ns, 256), 256, 0, st>>>(g, use_drift, w->code, use_drift ? w->drift : nullptr, w->act, w->na, 0, (size_t)g.nf * g.ns);
	FG_CUDA(cudaMemsetAsync(w->counter, 0, 4, st));
	// the completion flags are cleared per solve (a constant "done" value keeps the launch arguments identical from step to
	// step, so a captured whole-step CUDA graph stays valid)
	FG_CUDA(cudaMemsetAsync(w->flags, 0, (size_t)w->ni * w->nj * w->nq * 4, st));
	w->epoch = 1;
	TiledArgs A;
	A.nf = g.nf; A.ns = g.ns; A.vel_f = g.vel_f; A.vel_s = g.vel_s; A.p = g.p; A.code = w->code;
	A.drift = use_drift ? w->drift : nullptr; A.tiles = w->tiles; A.n_tiles = w->n_tiles; A.ni = w->ni; A.nj = w->nj;
	A.kc = w->kc; A.flags = w->flags; A.counter = w->counter; A.epoch = w->epoch; A.cp = cp; A.omega = omega; A.prof = w->prof;
	A.act = w->act; A.na = w->na; A.nb = w->nb; A.status = w->status; A.host_status = w->h_status;
	const int kw = w->kc <= 8 ? 8 : 16;  // warps per CTA: the box staging is unrolled for exactly this many threads
	dim3 block(kTile, kw);
	int per_sm = 0;
	const void* fn;
#define FG_TILED_FN(KWV)                                                                                                          \
	(g.x_fast ? (use_drift ? (const void*)sor_gs_tiled_kernel<true, true, KWV> : (const void*)sor_gs_tiled_kernel<true, false, KWV>) \
	          : (use_drift ? (const void*)sor_gs_tiled_kernel<false, true, KWV> : (const void*)sor_gs_tiled_kernel<false, false, KWV>))
	if (kw == 8) fn = FG_TILED_FN(8);
	else fn = FG_TILED_FN(16);
#undef FG_TILED_FN
	const size_t tile_smem = (size_t)kBoxW * kBoxW * 4 * (use_drift ? 5 : 4);
	FG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
	FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kTile * kw, tile_smem));
	if (const char* e = getenv("FLIPGPU_TILE_CTAS")) per_sm = std::min(per_sm, std::max(1, atoi(e)));  // tuning knob
	int blocks = std::min(w->n_tiles, std::max(1, per_sm) * w->sms);
	void* args[] = {&A};
	FG_CUDA(cudaLaunchKernel(fn, dim3(blocks), block, args, tile_smem, st));
		if (lc) lc->n += 2;
	return FLIPGPU_OK;
}


// ===================================================================================================
// Temporally blocked RED-BLACK SOR (north_star: "red-black SOR pressure solve ... staged through shared memory").
// A CTA stages a (T+2K+1)^2 box of u, v, p, the cell code and the drift term in shared memory, runs K colour passes on
// it -- the region whose inputs are still exact shrinks by one cell per pass on every side -- and writes the exact T x T
// interior to the OUTPUT arrays (ping-pong: a neighbouring tile may still be reading this tile's cells as its halo).
// Halo cells are recomputed by every tile that needs them, with identical arithmetic on identical inputs, so the
// result is bit-identical to one launch per colour (sor_rb_pass_kernel) at 1/K of the HBM sweeps; K = 8 passes also
// matches the 9-row deep halo of the slab-sharded solve (one NCCL exchange per launch).
constexpr int kRbT = 64, kRbK = 8, kRbW = kRbT + 2 * kRbK + 1;
// Box rows are stored de-interleaved by column parity (even columns first, then odd): the cells of one colour in a
// row, their x+1 neighbours and their y+1 neighbours are then each contiguous in shared memory -- no bank conflicts.
constexpr int kRbWh = (kRbW + 1) / 2, kRbRow = 2 * kRbWh, kRbBox = kRbRow * kRbW;
__device__ __forceinline__ int rb_index(int lx, int ly) { return ly * kRbRow + (lx & 1) * kRbWh + (lx >> 1); }

struct RbArgs {
	int nf, ns;
	const float *in_f, *in_s, *in_p;
	float *out_f, *out_s, *out_p;
	const unsigned char* code;
	const float* drift;
	int row_lo, row_hi;  // output rows [row_lo, row_hi)
	int passes, colour0; // colour passes in this launch (<= K) and the colour of the first one
	int tiles_x;
	float cp, omega;
};

template <bool XFAST, bool DRIFT>
__global__ void __launch_bounds__(512, 2) sor_rb_tiled_kernel(RbArgs A) {
	extern __shared__ float smem_rb[];
	float* s_f = smem_rb;
	float* s_s = s_f + kRbBox;
	float* s_p = s_s + kRbBox;
	float* s_d = s_p + kRbBox;  // only touched when DRIFT
	unsigned char* s_c = reinterpret_cast<unsigned char*>(s_d + (DRIFT ? kRbBox : 0));
	const int tid = threadIdx.y * 32 + threadIdx.x;
	const int X0 = (blockIdx.x % A.tiles_x) * kRbT, Y0 = A.row_lo + (blockIdx.x / A.tiles_x) * kRbT;
	const int BX = X0 - kRbK, BY = Y0 - kRbK;
	// stage the box: batches of 4 elements per thread with all loads of a batch in flight before any is stored
	for (int e0 = tid; e0 < kRbW * kRbW; e0 += 4 * 512) {
		float vf[4], vs[4], pp[4], dr[4];
		unsigned char cd[4];
#pragma unroll
		for (int r = 0; r < 4; r++) {
			int e = e0 + r * 512;
			int lx = e % kRbW, ly = e / kRbW;
			int F = BX + lx, S = BY + ly;
			vf[r] = 0.f; vs[r] = 0.f; pp[r] = 0.f; dr[r] = 0.f; cd[r] = 0;
			if (e < kRbW * kRbW && F >= 0 && F < A.nf && S >= 0 && S < A.ns) {
				size_t c = (size_t)F + (size_t)A.nf * S;
				vf[r] = A.in_f[c]; vs[r] = A.in_s[c]; pp[r] = A.in_p[c]; cd[r] = A.code[c];
				if (DRIFT) dr[r] = A.drift[c];
			}
		}
#pragma unroll
		for (int r = 0; r < 4; r++) {
			int e = e0 + r * 512;
			if (e < kRbW * kRbW) {
				int b = rb_index(e % kRbW, e / kRbW);
				s_f[b] = vf[r]; s_s[b] = vs[r]; s_p[b] = pp[r]; s_c[b] = cd[r];
				if (DRIFT) s_d[b] = dr[r];
			}
		}
	}
	__syncthreads();
	for (int m = 0; m < A.passes; m++) {
		const int colour = (A.colour0 + m) & 1, hi = kRbW - 2 - m;
		// a thread's rows are 16 apart, so they share the parity: its columns and box offsets are fixed for the pass and the
		// row loop only adds a constant
		const int ly0 = m + threadIdx.y;
		const int lxa = m + ((BX + m + BY + ly0 + colour) & 1) + 2 * threadIdx.x;
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const int lx = lxa + 64 * half;
			if (lx > hi) break;
			int e = rb_index(lx, ly0), ex = rb_index(lx + 1, ly0);
			for (int ly = ly0; ly <= hi; ly += 16, e += 16 * kRbRow, ex += 16 * kRbRow) {
				const int ey = e + kRbRow;
				const unsigned cd = s_c[e];
				if (!(cd & 1u)) continue;  // code 0 outside the grid interior, on solid/air cells and where s_sum == 0
				float vfc = s_f[e], vfp = s_f[ex], vsc = s_s[e], vsp = s_s[ey];
				float div = XFAST ? ((vfp - vfc) + vsp) - vsc : ((vsp - vsc) + vfp) - vfc;
				if (DRIFT) div -= s_d[e];
				if (cd == 31u) {
					float pv = (-div * 0.25f) * A.omega;
					s_p[e] += A.cp * pv;
					s_f[e] = vfc - pv; s_f[ex] = vfp + pv; s_s[e] = vsc - pv; s_s[ey] = vsp + pv;
				} else {
					float sfm = (cd & 2u) ? 1.f : 0.f, sfp = (cd & 4u) ? 1.f : 0.f;
					float ssm = (cd & 8u) ? 1.f : 0.f, ssp = (cd & 16u) ? 1.f : 0.f;
					float ssum = XFAST ? ((sfm + sfp) + ssm) + ssp : ((ssm + ssp) + sfm) + sfp;
					float pv = -div / ssum;
					pv *= A.omega;
					s_p[e] += A.cp * pv;
					s_f[e] = vfc - sfm * pv; s_f[ex] = vfp + sfp * pv;
					s_s[e] = vsc - ssm * pv; s_s[ey] = vsp + ssp * pv;
				}
			}
		}
		__syncthreads();
	}
	for (int e = tid; e < kRbT * kRbT; e += 512) {
		int ix = e % kRbT, iy = e / kRbT;
		int F = X0 + ix, S = Y0 + iy;
		if (F >= A.nf || S >= A.row_hi) continue;
		int b = rb_index(ix + kRbK, iy + kRbK);
		size_t c = (size_t)F + (size_t)A.nf * S;
		A.out_f[c] = s_f[b]; A.out_s[c] = s_s[b]; A.out_p[c] = s_p[b];
	}
}

// per-solve cell constants (code, drift) for the temporally blocked red-black kernel
int sor_rb_prep(SorWorkspace** wsp, const SorGrid& g, bool drift, int row_lo, int row_hi, cudaStream_t st, LaunchCounter* lc) {
	if (!*wsp) *wsp = new SorWorkspace;
	SorWorkspace* w = *wsp;
	FG_TRY(ensure_cell_constants(w, g));
	bool use_drift = drift && g.density && g.rest;
	const size_t c0 = (size_t)g.nf * std::max(row_lo, 0), c1 = (size_t)g.nf * std::min(row_hi, g.ns);
	sor_prep_kernel<<<wave_grid(c1 - c0, 256), 256, 0, st>>>(g, use_drift, w->code, use_drift ? w->drift : nullptr, nullptr, 0, c0, c1);
	if (lc) lc->n++;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

// `passes` (<= 8) colour passes, first colour = colour0, reading g's arrays and writing rows [row_lo,row_hi) of the out arrays
int sor_rb_round(SorWorkspace* w, const SorGrid& g, float* out_f, float* out_s, float* out_p, int passes, int colour0, int row_lo,
                 int row_hi, float cp, float omega, bool drift, cudaStream_t st, LaunchCounter* lc) {
	if (row_hi <= row_lo || passes <= 0) return FLIPGPU_OK;
	bool use_drift = drift && g.density && g.rest;
	RbArgs A;
	A.nf = g.nf; A.ns = g.ns; A.in_f = g.vel_f; A.in_s = g.vel_s; A.in_p = g.p; A.out_f = out_f; A.out_s = out_s; A.out_p = out_p;
	A.code = w->code; A.drift = use_drift ? w->drift : nullptr; A.row_lo = row_lo; A.row_hi = row_hi;
	A.passes = std::min(passes, kRbK); A.colour0 = colour0 & 1; A.tiles_x = (g.nf + kRbT - 1) / kRbT; A.cp = cp; A.omega = omega;
	const int tiles_y = (row_hi - row_lo + kRbT - 1) / kRbT;
	// (a variant that batched a thread's five cells per colour pass -- two shared-memory round trips per half pass -- measured
	// 4.58 vs 4.06 ms per solve at S2 on B200 and was dropped: profiles/r2_ab_v2_variants_4096.jsonl)
	const void* fn = g.x_fast ? (use_drift ? (const void*)sor_rb_tiled_kernel<true, true> : (const void*)sor_rb_tiled_kernel<true, false>)
	                          : (use_drift ? (const void*)sor_rb_tiled_kernel<false, true> : (const void*)sor_rb_tiled_kernel<false, false>);
	const size_t smem = (size_t)kRbBox * ((use_drift ? 4 : 3) * sizeof(float) + 1);
	FG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	void* args[] = {&A};
	FG_CUDA(cudaLaunchKernel(fn, dim3(A.tiles_x * tiles_y), dim3(32, 16), args, smem, st));
	if (lc) lc->n++;
	return FLIPGPU_OK;
}
int sor_rb_passes_per_round() { return kRbK; }
// A view whose rows are a WINDOW of a larger grid (the slab of sor_shard.cu) must not update the rows that are the
// larger grid's border or lie outside it: clear their cell codes after sor_rb_prep.  Rows [row_lo, row_hi) of the view.
int sor_rb_deactivate_rows(SorWorkspace* w, const SorGrid& g, int row_lo, int row_hi, cudaStream_t st) {
	row_lo = std::max(row_lo, 0); row_hi = std::min(row_hi, g.ns);
	if (!w || !w->code || row_hi <= row_lo) return FLIPGPU_OK;
	FG_CUDA(cudaMemsetAsync(w->code + (size_t)g.nf * row_lo, 0, (size_t)g.nf * (row_hi - row_lo), st));
	return FLIPGPU_OK;
}

}  // namespace fg
