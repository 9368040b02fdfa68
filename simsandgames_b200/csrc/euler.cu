// B200-native Eulerian smoke solver behind the C ABI of include/flipgpu.h.
// Replaces class Fluid (reference eulerian_fluid/src/fluid.h:19-260).  Storage is the reference's:
// ix(i,j) = i*num_y + j, y fastest (fluid.h:94-96), so threadIdx.x runs along j.  The projection is the
// shared SOR of sor.cu (s = !solid, cell_type = solid ? SOLID : FLUID, omega = 1.9 as fluid.h:127).
// Advection is ping-pong (new_u/new_v/new_m are written in full, then pointers swap) instead of the
// reference's copy-in / copy-out (fluid.h:212-213,237-238,242,258) -- same values, half the traffic.
// Compiled with -fmad=false: every fp32 expression keeps the reference's association order.
#include <algorithm>
#include <cmath>
#include "common.cuh"

namespace fg {

struct EGeo {
	int nx, ny;
	float h, h1, h2;
};
enum { U_FIELD = 0, V_FIELD = 1, S_FIELD = 2 };  // fluid.h:33-37

#define EIX(i, j) ((i) * g.ny + (j))

// sampleField, fluid.h:153-188
__device__ __forceinline__ float sample_field(const EGeo& g, const float* __restrict__ f, float x, float y, float dx, float dy) {
	x = fmaxf(g.h, fminf(x, (float)g.nx * g.h));
	y = fmaxf(g.h, fminf(y, (float)g.ny * g.h));
	int x0 = min((int)(g.h1 * (x - dx)), g.nx - 1);
	int y0 = min((int)(g.h1 * (y - dy)), g.ny - 1);
	int x1 = min(x0 + 1, g.nx - 1), y1 = min(y0 + 1, g.ny - 1);
	float tx = g.h1 * ((x - dx) - g.h * (float)x0);
	float ty = g.h1 * ((y - dy) - g.h * (float)y0);
	float sx = 1.f - tx, sy = 1.f - ty;
	return sx * sy * f[EIX(x0, y0)] + sx * ty * f[EIX(x0, y1)] + tx * sy * f[EIX(x1, y0)] + tx * ty * f[EIX(x1, y1)];
}

// advectVel, fluid.h:210-239 (avgU :191-198, avgV :201-208); writes every cell of new_u/new_v
__global__ void __launch_bounds__(256) k_advect_vel(EGeo g, float dt, const float* __restrict__ u, const float* __restrict__ v,
                                                    const unsigned char* __restrict__ solid, float* __restrict__ new_u,
                                                    float* __restrict__ new_v) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	int i = blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx || j >= g.ny) return;
	int c = EIX(i, j);
	float nu = u[c], nv = v[c];
	if (i >= 1 && j >= 1 && !solid[c]) {
		if (!solid[EIX(i - 1, j)] && j < g.ny - 1) {
			float x = g.h * (float)i, y = g.h2 + g.h * (float)j;
			float u0 = u[c];
			float v0 = (v[c] + v[EIX(i, j + 1)] + v[EIX(i - 1, j)] + v[EIX(i - 1, j + 1)]) / 4.f;
			nu = sample_field(g, u, x - dt * u0, y - dt * v0, 0.f, g.h2);
		}
		if (!solid[EIX(i, j - 1)] && i < g.nx - 1) {
			float x = g.h2 + g.h * (float)i, y = g.h * (float)j;
			float u0 = (u[c] + u[EIX(i, j - 1)] + u[EIX(i + 1, j)] + u[EIX(i + 1, j - 1)]) / 4.f;
			float v0 = v[c];
			nv = sample_field(g, v, x - dt * u0, y - dt * v0, g.h2, 0.f);
		}
	}
	new_u[c] = nu;
	new_v[c] = nv;
}

// advectSmoke, fluid.h:241-259
__global__ void __launch_bounds__(256) k_advect_smoke(EGeo g, float dt, const float* __restrict__ u, const float* __restrict__ v,
                                                      const unsigned char* __restrict__ solid, const float* __restrict__ m,
                                                      float* __restrict__ new_m) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	int i = blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx || j >= g.ny) return;
	int c = EIX(i, j);
	float nm = m[c];
	if (i >= 1 && i < g.nx - 1 && j >= 1 && j < g.ny - 1 && !solid[c]) {
		float u0 = (u[c] + u[EIX(i + 1, j)]) / 2.f;
		float v0 = (v[c] + v[EIX(i, j + 1)]) / 2.f;
		float x = g.h2 + g.h * (float)i - dt * u0;
		float y = g.h2 + g.h * (float)j - dt * v0;
		nm = sample_field(g, m, x, y, g.h2, g.h2);
	}
	new_m[c] = nm;
}

// extrapolate, fluid.h:142-151.  The two loops touch disjoint arrays; within each, the reference's
// serial order only matters where a row/column has fewer than 3 cells, which the ctor excludes (x,y>=1).
__global__ void k_extrapolate(EGeo g, float* __restrict__ u, float* __restrict__ v) {
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < g.nx) {
		u[EIX(t, 0)] = u[EIX(t, 1)];
		u[EIX(t, g.ny - 1)] = u[EIX(t, g.ny - 2)];
	}
	if (t < g.ny) {
		v[EIX(0, t)] = v[EIX(1, t)];
		v[EIX(g.nx - 1, t)] = v[EIX(g.nx - 2, t)];
	}
}

// FluidUI::setObstacle, eulerian_fluid/src/fluid_ui.h:24-62: clear `solid`, left/top/bottom walls, then a solid disc
// that carries (vx, vy) and resets its smoke to 1.  One thread per cell instead of memset + three loops; the disc test
// is the reference's int arithmetic (dx*dx + dy*dy < r*r).
__global__ void __launch_bounds__(256) k_euler_obstacle(EGeo g, int ox, int oy, int r, float vx, float vy,
                                                        unsigned char* __restrict__ solid, float* __restrict__ m,
                                                        float* __restrict__ u, float* __restrict__ v) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	int i = blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx || j >= g.ny) return;
	int c = EIX(i, j);
	bool sd = i == 0 || j == 0 || j == g.ny - 1;
	if (i >= 1 && i < g.nx - 1 && j >= 1 && j < g.ny - 1) {
		int dx = i - ox, dy = j - oy;
		if (dx * dx + dy * dy < r * r) {
			sd = true;
			m[c] = 1.f;
			u[c] = vx;
			v[c] = vy;
		}
	}
	solid[c] = sd ? 1 : 0;
}

__global__ void k_solid_to_coeff(int n, const unsigned char* __restrict__ solid, float* __restrict__ s, int* __restrict__ cell_type) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
		bool sd = solid[c] != 0;
		s[c] = sd ? 0.f : 1.f;
		cell_type[c] = sd ? SOLID_CELL : FLUID_CELL;
	}
}
__global__ void k_fill_f32(int n, float* a, float val) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) a[c] = val;
}
__global__ void k_sample_points(EGeo g, const float* __restrict__ f, float dx, float dy, const float* __restrict__ xs,
                                const float* __restrict__ ys, float* __restrict__ out, size_t n) {
	for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
		out[k] = sample_field(g, f, xs[k], ys[k], dx, dy);
}

}  // namespace fg

using namespace fg;

struct EulerSim {
	EulerDims d{};
	int device = 0;
	cudaStream_t st = nullptr;
	float *u = nullptr, *v = nullptr, *new_u = nullptr, *new_v = nullptr, *pressure = nullptr, *m = nullptr, *new_m = nullptr;
	unsigned char* solid = nullptr;
	float *u_alt = nullptr, *v_alt = nullptr, *p_alt = nullptr;  // ping-pong partners for the tiled red-black projection (lazy)
	float* s = nullptr;        // derived: !solid
	int* cell_type = nullptr;  // derived
	bool coeff_dirty = true;
	float* d_partials = nullptr;
	SorWorkspace* ws = nullptr;
	LaunchCounter lc;
};

namespace {
constexpr int kEPartials = 1024;
EGeo egeo(const EulerSim* e) {
	EGeo g;
	g.nx = e->d.num_x; g.ny = e->d.num_y; g.h = e->d.h; g.h1 = e->d.h1; g.h2 = e->d.h / 2;
	return g;
}
struct EDeviceGuard {
	int prev = -1;
	explicit EDeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
	~EDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
int refresh_coeff(EulerSim* e) {
	if (!e->coeff_dirty) return FLIPGPU_OK;
	k_solid_to_coeff<<<wave_grid(e->d.num_cells, 256), 256, 0, e->st>>>(e->d.num_cells, e->solid, e->s, e->cell_type);
	e->lc.n++;
	FG_CUDA(cudaGetLastError());
	e->coeff_dirty = false;
	return FLIPGPU_OK;
}
SorGrid esor(EulerSim* e) {
	SorGrid g;
	g.nf = e->d.num_y; g.ns = e->d.num_x;
	g.vel_f = e->v; g.vel_s = e->u; g.p = e->pressure; g.s = e->s; g.cell_type = e->cell_type;
	g.density = nullptr; g.rest = nullptr; g.x_fast = false;
	g.alt_f = e->v_alt; g.alt_s = e->u_alt; g.alt_p = e->p_alt;
	return g;
}
float** efield(EulerSim* e, int field) {
	switch (field) {
		case EULER_U: return &e->u; case EULER_V: return &e->v; case EULER_NEW_U: return &e->new_u;
		case EULER_NEW_V: return &e->new_v; case EULER_PRESSURE: return &e->pressure; case EULER_M: return &e->m;
		case EULER_NEW_M: return &e->new_m;
	}
	return nullptr;
}
int project(EulerSim* e, int iters, float dt, int order) {
	FG_TRY(refresh_coeff(e));
	FG_CUDA(cudaMemsetAsync(e->pressure, 0, (size_t)e->d.num_cells * 4, e->st));  // fluid.h:101-103
	float cp = e->d.density * e->d.h / dt;                                          // fluid.h:100
	if (order == FLIP_SOLVER_RED_BLACK && !e->u_alt) {
		size_t C = e->d.num_cells;
		FG_CUDA(cudaMalloc(&e->u_alt, C * 4)); FG_CUDA(cudaMalloc(&e->v_alt, C * 4)); FG_CUDA(cudaMalloc(&e->p_alt, C * 4));
	}
	bool swapped = false;
	FG_TRY(sor_solve(&e->ws, esor(e), iters, cp, 1.9f, false, order, true, e->st, &e->lc, &swapped));
	if (swapped) { std::swap(e->u, e->u_alt); std::swap(e->v, e->v_alt); std::swap(e->pressure, e->p_alt); }
	return FLIPGPU_OK;
}
int extrapolate(EulerSim* e) {
	EGeo g = egeo(e);
	int n = std::max(g.nx, g.ny);
	k_extrapolate<<<cdiv(n, 256), 256, 0, e->st>>>(g, e->u, e->v);
	e->lc.n++;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}
int advect_vel(EulerSim* e, float dt) {
	EGeo g = egeo(e);
	dim3 block(64, 4), grid(cdiv(g.ny, 64), cdiv(g.nx, 4));
	k_advect_vel<<<grid, block, 0, e->st>>>(g, dt, e->u, e->v, e->solid, e->new_u, e->new_v);
	e->lc.n++;
	FG_CUDA(cudaGetLastError());
	std::swap(e->u, e->new_u);  // u <- new_u ; the old field stays readable as new_u (the reference leaves a copy there)
	std::swap(e->v, e->new_v);
	return FLIPGPU_OK;
}
int advect_smoke(EulerSim* e, float dt) {
	EGeo g = egeo(e);
	dim3 block(64, 4), grid(cdiv(g.ny, 64), cdiv(g.nx, 4));
	k_advect_smoke<<<grid, block, 0, e->st>>>(g, dt, e->u, e->v, e->solid, e->m, e->new_m);
	e->lc.n++;
	FG_CUDA(cudaGetLastError());
	std::swap(e->m, e->new_m);
	return FLIPGPU_OK;
}
}  // namespace

extern "C" {

int euler_create(int x, int y, float density, float h, int device, EulerSim** out) {
	FG_CHECK(out && x >= 1 && y >= 1 && h > 0, FLIPGPU_EINVAL, "euler_create: bad argument");
	FG_CHECK((long long)(x + 2) * (y + 2) < (1ll << 31), FLIPGPU_EINVAL, "euler_create: grid too large for int32");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		fg::set_error("euler_create: no CUDA device (this library has no CPU fallback)");
		return FLIPGPU_ENODEV;
	}
	if (device < 0) FG_CUDA(cudaGetDevice(&device));
	FG_CHECK(device < ndev, FLIPGPU_EINVAL, "euler_create: device %d of %d", device, ndev);
	EulerSim* e = new EulerSim;
	e->device = device;
	EDeviceGuard guard(device);
	e->d.num_x = 2 + x; e->d.num_y = 2 + y; e->d.num_cells = e->d.num_x * e->d.num_y;  // fluid.h:43-44
	e->d.density = density; e->d.h = h; e->d.h1 = 1 / h;
	FG_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
	size_t C = e->d.num_cells;
	float** fs[] = {&e->u, &e->v, &e->new_u, &e->new_v, &e->pressure, &e->m, &e->new_m, &e->s};
	for (float** fp : fs) {
		FG_CUDA(cudaMalloc(fp, C * 4));
		FG_CUDA(cudaMemsetAsync(*fp, 0, C * 4, e->st));
	}
	FG_CUDA(cudaMalloc(&e->solid, C));
	FG_CUDA(cudaMemsetAsync(e->solid, 0, C, e->st));
	FG_CUDA(cudaMalloc(&e->cell_type, C * 4));
	FG_CUDA(cudaMalloc(&e->d_partials, (3 * kEPartials + 8) * 4));
	k_fill_f32<<<wave_grid(C, 256), 256, 0, e->st>>>((int)C, e->m, 1.f);  // fluid.h:63-68
	e->lc.n++;
	FG_CUDA(cudaStreamSynchronize(e->st));
	*out = e;
	return FLIPGPU_OK;
}

int euler_destroy(EulerSim* e) {
	if (!e) return FLIPGPU_OK;
	EDeviceGuard guard(e->device);
	cudaStreamSynchronize(e->st);
	void* ptrs[] = {e->u, e->v, e->new_u, e->new_v, e->pressure, e->m, e->new_m, e->s, e->solid, e->cell_type, e->d_partials, e->u_alt, e->v_alt, e->p_alt};
	for (void* p : ptrs) if (p) cudaFree(p);
	sor_workspace_free(e->ws);
	cudaStreamDestroy(e->st);
	delete e;
	return FLIPGPU_OK;
}

int euler_get_dims(EulerSim* e, EulerDims* dims) {
	FG_CHECK(e && dims, FLIPGPU_EINVAL, "null argument");
	*dims = e->d;
	return FLIPGPU_OK;
}

int euler_upload(EulerSim* e, int field, const void* host, size_t count) {
	FG_CHECK(e && host, FLIPGPU_EINVAL, "euler_upload: null argument");
	FG_CHECK(field >= 0 && field < EULER_FIELD_COUNT, FLIPGPU_EINVAL, "euler_upload: unknown field %d", field);
	FG_CHECK(count == (size_t)e->d.num_cells, FLIPGPU_EINVAL, "euler_upload: expected %d elements, got %zu", e->d.num_cells, count);
	EDeviceGuard guard(e->device);
	if (field == EULER_SOLID) {
		FG_CUDA(cudaMemcpyAsync(e->solid, host, count, cudaMemcpyHostToDevice, e->st));
		e->coeff_dirty = true;
	} else {
		FG_CUDA(cudaMemcpyAsync(*efield(e, field), host, count * 4, cudaMemcpyHostToDevice, e->st));
	}
	FG_CUDA(cudaStreamSynchronize(e->st));
	return FLIPGPU_OK;
}

int euler_readback(EulerSim* e, int field, void* host, size_t count) {
	FG_CHECK(e && host, FLIPGPU_EINVAL, "euler_readback: null argument");
	FG_CHECK(field >= 0 && field < EULER_FIELD_COUNT, FLIPGPU_EINVAL, "euler_readback: unknown field %d", field);
	FG_CHECK(count == (size_t)e->d.num_cells, FLIPGPU_EINVAL, "euler_readback: expected %d elements, got %zu", e->d.num_cells, count);
	EDeviceGuard guard(e->device);
	if (field == EULER_SOLID) FG_CUDA(cudaMemcpyAsync(host, e->solid, count, cudaMemcpyDeviceToHost, e->st));
	else FG_CUDA(cudaMemcpyAsync(host, *efield(e, field), count * 4, cudaMemcpyDeviceToHost, e->st));
	FG_CUDA(cudaStreamSynchronize(e->st));
	return sor_workspace_check(e->ws);
}

int euler_project(EulerSim* e, int iters, float dt, int order) {
	FG_CHECK(e && iters >= 0 && dt > 0, FLIPGPU_EINVAL, "euler_project: bad argument");
	EDeviceGuard guard(e->device);
	return project(e, iters, dt, order);
}
int euler_extrapolate(EulerSim* e) {
	FG_CHECK(e, FLIPGPU_EINVAL, "null handle");
	EDeviceGuard guard(e->device);
	return extrapolate(e);
}
int euler_advect_vel(EulerSim* e, float dt) {
	FG_CHECK(e, FLIPGPU_EINVAL, "null handle");
	EDeviceGuard guard(e->device);
	return advect_vel(e, dt);
}
int euler_advect_smoke(EulerSim* e, float dt) {
	FG_CHECK(e, FLIPGPU_EINVAL, "null handle");
	EDeviceGuard guard(e->device);
	return advect_smoke(e, dt);
}
int euler_step(EulerSim* e, int iters, float dt, int order, int n_steps) {
	FG_CHECK(e && iters >= 0 && dt > 0, FLIPGPU_EINVAL, "euler_step: bad argument");
	EDeviceGuard guard(e->device);
	for (int k = 0; k < n_steps; k++) {  // eulerian_fluid/src/fluid_ui.h:107-111
		FG_TRY(project(e, iters, dt, order));
		FG_TRY(extrapolate(e));
		FG_TRY(advect_vel(e, dt));
		FG_TRY(advect_smoke(e, dt));
	}
	return FLIPGPU_OK;
}

int euler_set_obstacle(EulerSim* e, int x, int y, int r, float vx, float vy) {
	FG_CHECK(e, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(r >= 0 && r < 46341, FLIPGPU_EINVAL, "euler_set_obstacle: radius %d", r);
	EDeviceGuard guard(e->device);
	EGeo g = egeo(e);
	dim3 block(64, 4), grid(cdiv(g.ny, 64), cdiv(g.nx, 4));
	k_euler_obstacle<<<grid, block, 0, e->st>>>(g, x, y, r, vx, vy, e->solid, e->m, e->u, e->v);
	e->lc.n++;
	FG_CUDA(cudaGetLastError());
	e->coeff_dirty = true;
	return FLIPGPU_OK;
}

int euler_sample_field(EulerSim* e, int field, const float* xs, const float* ys, float* out, size_t n) {
	FG_CHECK(e && xs && ys && out, FLIPGPU_EINVAL, "euler_sample_field: null argument");
	FG_CHECK(field >= U_FIELD && field <= S_FIELD, FLIPGPU_EINVAL, "euler_sample_field: field %d", field);
	if (n == 0) return FLIPGPU_OK;
	EDeviceGuard guard(e->device);
	EGeo g = egeo(e);
	float* d = nullptr;
	FG_CUDA(cudaMalloc(&d, 3 * n * 4));
	FG_CUDA(cudaMemcpyAsync(d, xs, n * 4, cudaMemcpyHostToDevice, e->st));
	FG_CUDA(cudaMemcpyAsync(d + n, ys, n * 4, cudaMemcpyHostToDevice, e->st));
	const float* f = field == U_FIELD ? e->u : (field == V_FIELD ? e->v : e->m);
	float dx = field == U_FIELD ? 0.f : g.h2, dy = field == V_FIELD ? 0.f : g.h2;
	k_sample_points<<<wave_grid(n, 256), 256, 0, e->st>>>(g, f, dx, dy, d, d + n, d + 2 * n, n);
	e->lc.n++;
	cudaError_t err = cudaMemcpyAsync(out, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, e->st);
	if (err == cudaSuccess) err = cudaStreamSynchronize(e->st);
	cudaFree(d);
	FG_CUDA(err);
	return FLIPGPU_OK;
}

int euler_get_residual(EulerSim* e, float* max_div, float* rms_div) {
	FG_CHECK(e && max_div && rms_div, FLIPGPU_EINVAL, "null argument");
	EDeviceGuard guard(e->device);
	FG_TRY(refresh_coeff(e));
	float r2[2];
	FG_TRY(sor_residual(esor(e), false, e->d_partials, kEPartials, r2, e->st, &e->lc));
	*max_div = r2[0]; *rms_div = r2[1];
	return FLIPGPU_OK;
}
int euler_synchronize(EulerSim* e) {
	FG_CHECK(e, FLIPGPU_EINVAL, "null handle");
	EDeviceGuard guard(e->device);
	FG_CUDA(cudaStreamSynchronize(e->st));
	return sor_workspace_check(e->ws);
}
int euler_get_stream(EulerSim* e, void** stream) {
	FG_CHECK(e && stream, FLIPGPU_EINVAL, "null argument");
	*stream = (void*)e->st;
	return FLIPGPU_OK;
}
int euler_get_launch_count(EulerSim* e, long long* launches) {
	FG_CHECK(e && launches, FLIPGPU_EINVAL, "null argument");
	*launches = e->lc.n;
	return FLIPGPU_OK;
}

}  // extern "C"
