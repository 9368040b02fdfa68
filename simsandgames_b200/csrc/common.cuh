// Shared plumbing of libflipgpu.so: error reporting, launch accounting, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include "../../include/flipgpu.h"

namespace fg {

void set_error(const char* fmt, ...);

#define FG_CUDA(call)                                                                            \
	do {                                                                                         \
		cudaError_t e_ = (call);                                                                 \
		if (e_ != cudaSuccess) {                                                                 \
			fg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return FLIPGPU_ECUDA;                                                                \
		}                                                                                        \
	} while (0)

#define FG_CHECK(cond, code, ...)       \
	do {                                \
		if (!(cond)) {                  \
			fg::set_error(__VA_ARGS__); \
			return code;                \
		}                               \
	} while (0)

#define FG_TRY(expr)              \
	do {                          \
		int rc_ = (expr);         \
		if (rc_ != FLIPGPU_OK) return rc_; \
	} while (0)

// cell types of flip_fluid.h:30-34
enum : int { FLUID_CELL = 0, AIR_CELL = 1, SOLID_CELL = 2 };

// SM count of the current device (148 on B200), queried once
inline int num_sms() {
	static int n = 0;
	if (n == 0) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
	}
	return n;
}

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Grid size for grid-stride kernels: whole waves of the SMs, capped by the work available.  Eight waves rather than one resident
// wave: the CTAs of a single wave start together and run their memory-stalled and their arithmetic stretches in step, a
// stream of shorter CTAs interleaves them (separation sweeps at S2: 3.96 -> 3.33 ms, whole step 12.65 -> 11.91 ms; 64 waves: no further gain).
constexpr int kWavesPerGrid = 8;
inline unsigned wave_grid(size_t work_items, int threads, int ctas_per_sm = 8) {
	size_t need = (work_items + threads - 1) / threads;
	size_t full = (size_t)num_sms() * ctas_per_sm * kWavesPerGrid;
	if (need >= full) return (unsigned)full;
	return (unsigned)(need ? need : 1);
}

struct LaunchCounter {
	long long n = 0;
};

// Views used by the shared SOR projection (sor.cuh).  "fast" is the contiguous axis of the
// storage: x for FlipFluid (fIX, flip_fluid.h:147-149), y for the Eulerian Fluid (ix, fluid.h:94-96).
struct SorGrid {
	int nf, ns;            // cells along the fast / slow axis
	float* vel_f;          // velocity component along the fast axis, at the low-index face (u for FLIP, v for Euler)
	float* vel_s;          // component along the slow axis (v for FLIP, u for Euler)
	float* p;              // pressure accumulator
	const float* s;        // 0 = solid, else open (flip_fluid.h:28)
	const int* cell_type;  // only FLUID_CELL cells are updated (flip_fluid.h:462 / fluid.h:110)
	const float* density;  // particle_density for the drift term, or nullptr
	const float* rest;     // device scalar particle_rest_density, or nullptr
	float *alt_f = nullptr, *alt_s = nullptr, *alt_p = nullptr;  // ping-pong partners of vel_f/vel_s/p for the tiled red-black
	                       // kernel (optional; without them the one-launch-per-colour kernel runs)
	bool x_fast;           // true: fast axis is x (FLIP); false: fast axis is y (Euler).  Fixes the fp32
	                       // association order of the divergence (flip_fluid.h:476 / fluid.h:122-123)
};

struct SorWorkspace;  // plan + per-cell constants of the blocked-wavefront solver (sor_tiled.cu)
void sor_workspace_free(SorWorkspace* w);
int sor_workspace_profile(SorWorkspace* w, unsigned long long out[8]);
int sor_workspace_check(SorWorkspace* w);  // after a stream sync: FLIPGPU_ESTATE if a solver dependency wait timed out
int sor_solve_tiled(SorWorkspace** ws, const SorGrid& g, int iters, float cp, float omega, bool drift, cudaStream_t st,
                    LaunchCounter* lc);
struct SysWorkspace;  // register-systolic wavefront kernel (sor_systolic.cu)
void sys_workspace_free(SysWorkspace* w);
int sys_workspace_check(SysWorkspace* w);
int sys_workspace_profile(SysWorkspace* w, unsigned long long out[8]);
SysWorkspace** sor_workspace_sys(SorWorkspace** ws);
int sor_solve_systolic(SysWorkspace** ws, const SorGrid& g, int iters, float cp, float omega, bool drift, cudaStream_t st,
                       LaunchCounter* lc);
// order: FlipSolverOrder.  s_binary: every s is 0 or 1 (required by the blocked kernel's 1-byte cell code;
// otherwise the lexicographic orders fall back to the unblocked diagonal kernel, same results).
int sor_solve(SorWorkspace** ws, const SorGrid& g, int iters, float cp, float omega, bool drift, int order, bool s_binary,
              cudaStream_t st, LaunchCounter* lc, bool* swapped = nullptr);
// temporally blocked red-black (sor_tiled.cu): cell constants once per solve, then rounds of up to 8 colour passes
int sor_rb_prep(SorWorkspace** ws, const SorGrid& g, bool drift, int row_lo, int row_hi, cudaStream_t st, LaunchCounter* lc);
int sor_rb_round(SorWorkspace* w, const SorGrid& g, float* out_f, float* out_s, float* out_p, int passes, int colour0, int row_lo,
                 int row_hi, float cp, float omega, bool drift, cudaStream_t st, LaunchCounter* lc);
int sor_rb_passes_per_round();
int sor_rb_deactivate_rows(SorWorkspace* w, const SorGrid& g, int row_lo, int row_hi, cudaStream_t st);
void sor_rb_rows_launch(const SorGrid& g, int colour, int shift, int lo, int hi, float cp, float omega, bool drift, cudaStream_t st);
int sor_residual(const SorGrid& g, bool drift, float* d_partials, int n_partials, float* h_out2, cudaStream_t st,
                 LaunchCounter* lc);

}  // namespace fg
