// SOR pressure projection shared by FlipFluid::solveIncompressibility (flip_fluid.h:451-495)
// and Fluid::solveIncompressibility (eulerian_fluid/src/fluid.h:98-139).
//
// One cell update (flip_fluid.h:461-491), identical in both sims once the Eulerian bool solid[] is
// expressed as s = !solid and cell_type = solid ? SOLID : FLUID:
//     skip unless cell_type==FLUID ; s_sum = s[l]+s[r]+s[b]+s[t] ; skip if 0
//     div = u[r]-u[c]+v[t]-v[c]  (- max(0, density-rest) when drift compensation is on)
//     pv = -div/s_sum*omega ; p[c]+=cp*pv ; u[c]-=s[l]pv ; u[r]+=s[r]pv ; v[c]-=s[b]pv ; v[t]+=s[t]pv
// A cell shares faces only with its 4 edge neighbours, so
//   * LEX_WAVEFRONT: sweeping anti-diagonals (i+j = const) in ascending order, with iteration k+1
//     trailing iteration k by two diagonals, performs exactly the reference's loop-order updates
//     -> bit-identical u,v,p (this file is compiled with -fmad=false);
//   * RED_BLACK: cells with (i+j) even, then odd: the performance order named by north_star.
#include <cooperative_groups.h>
#include <algorithm>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace fg {

template <bool XFAST, bool COHERENT>
__device__ __forceinline__ void sor_cell(const SorGrid& g, int c, float cp, float omega, float rest, bool drift) {
	const int nf = g.nf;
	if (g.cell_type[c] != FLUID_CELL) return;
	float sfm = g.s[c - 1], sfp = g.s[c + 1], ssm = g.s[c - nf], ssp = g.s[c + nf];
	float ssum = XFAST ? ((sfm + sfp) + ssm) + ssp : ((ssm + ssp) + sfm) + sfp;
	if (ssum == 0.f) return;
	float vfc, vfp, vsc, vsp;
	if (COHERENT) {  // wavefront kernel: values are produced by other CTAs between grid syncs
		vfc = __ldcg(g.vel_f + c); vfp = __ldcg(g.vel_f + c + 1);
		vsc = __ldcg(g.vel_s + c); vsp = __ldcg(g.vel_s + c + nf);
	} else {
		vfc = g.vel_f[c]; vfp = g.vel_f[c + 1];
		vsc = g.vel_s[c]; vsp = g.vel_s[c + nf];
	}
	float div = XFAST ? ((vfp - vfc) + vsp) - vsc : ((vsp - vsc) + vfp) - vfc;
	if (drift && rest > 0.f) {
		float compression = g.density[c] - rest;
		if (compression > 0.f) div -= compression;  // k = 1 (flip_fluid.h:479)
	}
	float pv = -div / ssum;
	pv *= omega;
	if (COHERENT) {
		__stcg(g.p + c, __ldcg(g.p + c) + cp * pv);
		__stcg(g.vel_f + c, vfc - sfm * pv);
		__stcg(g.vel_f + c + 1, vfp + sfp * pv);
		__stcg(g.vel_s + c, vsc - ssm * pv);
		__stcg(g.vel_s + c + nf, vsp + ssp * pv);
	} else {
		g.p[c] += cp * pv;
		g.vel_f[c] = vfc - sfm * pv;
		g.vel_f[c + 1] = vfp + sfp * pv;
		g.vel_s[c] = vsc - ssm * pv;
		g.vel_s[c + nf] = vsp + ssp * pv;
	}
}

// ---- red-black pass, one colour per launch (baseline kernel; the tiled multi-pass kernel is sor_tiled.cu)
template <bool XFAST>
__global__ void __launch_bounds__(256) sor_rb_pass_kernel(SorGrid g, int colour, float cp, float omega, bool drift) {
	int b = 1 + blockIdx.y * blockDim.y + threadIdx.y;
	if (b > g.ns - 2) return;
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	int a = 2 * k + ((b + colour) & 1);
	if (a < 1 || a > g.nf - 2) return;
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	sor_cell<XFAST, false>(g, a + g.nf * b, cp, omega, rest, drift);
}

// ---- lexicographic order as a pipelined anti-diagonal wavefront (cooperative launch)
template <bool XFAST>
__global__ void __launch_bounds__(256) sor_wavefront_kernel(SorGrid g, int iters, float cp, float omega, bool drift) {
	cg::grid_group grid = cg::this_grid();
	const int dmin = 2, dmax = (g.nf - 2) + (g.ns - 2);
	const int steps = (dmax - dmin + 1) + 2 * (iters - 1);
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	for (int t = 0; t < steps; t++) {
		for (int k = blockIdx.x; k < iters; k += gridDim.x) {
			int d = dmin + t - 2 * k;
			if (d < dmin || d > dmax) continue;
			int alo = max(1, d - (g.ns - 2)), ahi = min(g.nf - 2, d - 1);
			for (int a = alo + threadIdx.x; a <= ahi; a += blockDim.x)
				sor_cell<XFAST, true>(g, a + g.nf * (d - a), cp, omega, rest, drift);
		}
		grid.sync();
	}
}

int sor_solve(SorWorkspace** ws, const SorGrid& g, int iters, float cp, float omega, bool drift, int order, bool s_binary,
              cudaStream_t st, LaunchCounter* lc, bool* swapped) {
	if (swapped) *swapped = false;
	if (iters <= 0 || g.nf < 3 || g.ns < 3) return FLIPGPU_OK;
	bool use_drift = drift && g.density && g.rest;
	if (order == FLIP_SOLVER_LEX_WAVEFRONT && s_binary && ws) return sor_solve_tiled(ws, g, iters, cp, omega, drift, st, lc);
	if (order == FLIP_SOLVER_LEX_SYSTOLIC && s_binary && ws) return sor_solve_systolic(sor_workspace_sys(ws), g, iters, cp, omega, drift, st, lc);
	if (order == FLIP_SOLVER_LEX_WAVEFRONT || order == FLIP_SOLVER_LEX_DIAGONAL || order == FLIP_SOLVER_LEX_SYSTOLIC) {
		int dev = 0, sms = 0, per_sm = 0;
		FG_CUDA(cudaGetDevice(&dev));
		FG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		const void* fn = g.x_fast ? (const void*)sor_wavefront_kernel<true> : (const void*)sor_wavefront_kernel<false>;
		FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
		int blocks = iters < sms * per_sm ? iters : sms * per_sm;
		SorGrid gg = g;
		void* args[] = {&gg, &iters, &cp, &omega, &use_drift};
		FG_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, st));
		if (lc) lc->n++;
		return FLIPGPU_OK;
	}
	FG_CHECK(order == FLIP_SOLVER_RED_BLACK, FLIPGPU_EINVAL, "unknown solver order %d", order);
	if (s_binary && ws && swapped && g.alt_f && g.alt_s && g.alt_p) {
		// temporally blocked: 8 colour passes per launch, ping-pong between the arrays and their partners
		FG_TRY(sor_rb_prep(ws, g, drift, 0, g.ns, st, lc));
		SorGrid cur = g;
		float *of = g.alt_f, *os = g.alt_s, *op = g.alt_p;
		const int K = sor_rb_passes_per_round();
		for (int pass0 = 0; pass0 < 2 * iters; pass0 += K) {
			FG_TRY(sor_rb_round(*ws, cur, of, os, op, std::min(K, 2 * iters - pass0), pass0 & 1, 0, g.ns, cp, omega, drift, st, lc));
			std::swap(cur.vel_f, of); std::swap(cur.vel_s, os); std::swap(cur.p, op);
			*swapped = !*swapped;
		}
		return FLIPGPU_OK;
	}
	dim3 block(64, 4);
	dim3 grid(cdiv((size_t)(g.nf + 1) / 2 + 1, block.x), cdiv((size_t)g.ns - 2, block.y));
	for (int it = 0; it < iters; it++)
		for (int colour = 0; colour < 2; colour++) {
			if (g.x_fast) sor_rb_pass_kernel<true><<<grid, block, 0, st>>>(g, colour, cp, omega, use_drift);
			else sor_rb_pass_kernel<false><<<grid, block, 0, st>>>(g, colour, cp, omega, use_drift);
			if (lc) lc->n++;
		}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

// ---- residual: max |div| and sum div^2 over the cells the solver updates (drift term included)
template <bool XFAST>
__global__ void __launch_bounds__(256) sor_residual_kernel(SorGrid g, bool drift, float* partials) {
	float rest = (drift && g.rest) ? *g.rest : 0.f;
	float vmax = 0.f, vsq = 0.f, cnt = 0.f;
	size_t interior = (size_t)(g.nf - 2) * (g.ns - 2);
	for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < interior; idx += (size_t)gridDim.x * blockDim.x) {
		int a = 1 + (int)(idx % (g.nf - 2)), b = 1 + (int)(idx / (g.nf - 2));
		int c = a + g.nf * b;
		if (g.cell_type[c] != FLUID_CELL) continue;
		float ssum = ((g.s[c - 1] + g.s[c + 1]) + g.s[c - g.nf]) + g.s[c + g.nf];
		if (ssum == 0.f) continue;
		float div = XFAST ? ((g.vel_f[c + 1] - g.vel_f[c]) + g.vel_s[c + g.nf]) - g.vel_s[c]
		                  : ((g.vel_s[c + g.nf] - g.vel_s[c]) + g.vel_f[c + 1]) - g.vel_f[c];
		if (drift && rest > 0.f) {
			float compression = g.density[c] - rest;
			if (compression > 0.f) div -= compression;
		}
		vmax = fmaxf(vmax, fabsf(div));
		vsq += div * div;
		cnt += 1.f;
	}
	__shared__ float sm[3][8];
	for (int o = 16; o > 0; o >>= 1) {  // warp-shuffle reduction of the residual
		vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
		vsq += __shfl_xor_sync(0xffffffffu, vsq, o);
		cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	}
	int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { sm[0][w] = vmax; sm[1][w] = vsq; sm[2][w] = cnt; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int i = 1; i < 8; i++) { vmax = fmaxf(vmax, sm[0][i]); vsq += sm[1][i]; cnt += sm[2][i]; }
		partials[3 * blockIdx.x] = vmax; partials[3 * blockIdx.x + 1] = vsq; partials[3 * blockIdx.x + 2] = cnt;
	}
}

__global__ void sor_residual_final_kernel(const float* partials, int n, float* out2) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	float vmax = 0.f;
	double sq = 0.0, cnt = 0.0;
	for (int i = 0; i < n; i++) { vmax = fmaxf(vmax, partials[3 * i]); sq += partials[3 * i + 1]; cnt += partials[3 * i + 2]; }
	out2[0] = vmax;
	out2[1] = cnt > 0 ? (float)sqrt(sq / cnt) : 0.f;
}

// d_partials: at least 3*n_partials+2 floats of device scratch
int sor_residual(const SorGrid& g, bool drift, float* d_partials, int n_partials, float* h_out2, cudaStream_t st,
                 LaunchCounter* lc) {
	bool use_drift = drift && g.density && g.rest;
	int blocks = n_partials;
	if (g.x_fast) sor_residual_kernel<true><<<blocks, 256, 0, st>>>(g, use_drift, d_partials);
	else sor_residual_kernel<false><<<blocks, 256, 0, st>>>(g, use_drift, d_partials);
	sor_residual_final_kernel<<<1, 32, 0, st>>>(d_partials, blocks, d_partials + 3 * blocks);
	if (lc) lc->n += 2;
	FG_CUDA(cudaGetLastError());
	FG_CUDA(cudaMemcpyAsync(h_out2, d_partials + 3 * blocks, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
	FG_CUDA(cudaStreamSynchronize(st));
	return FLIPGPU_OK;
}

}  // namespace fg
