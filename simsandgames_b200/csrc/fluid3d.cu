// B200-native 3D Eulerian smoke step behind the C ABI of include/flipgpu.h.
// Replaces struct Fluid3D (reference eulerian_fluid/src/fluid3d.h:19-368, which no reference target compiles -- it is the
// 3D transcription of fluid.h "following the patterns from the 2d version", fluid3d.h:1).  Same state arrays, x-fastest
// layout ix(i,j,k) = i + num_x*j + num_y*num_x*k (:131-133), same four calls.  Compiled with -fmad=false: every kernel keeps the
// reference's unfused fp32 arithmetic and association order, so each call is bit-identical to the serial code.
//
// Reference behaviour reproduced on purpose (quirk register):
//   Q19  advectVel's k loop runs to num_y, not num_z (:311): layers k >= num_y keep their velocity.  For num_z < num_y the
//        reference reads out of bounds (undefined); here the loop stops at min(num_y, num_z).
//   Q20  sampleField(S_FIELD) offsets x and y by h/2 but not z (:247).
//   Q21  extrapolate() runs its three border loops one after the other (:186-210): edge cells see the earlier loops' values.
#include <cooperative_groups.h>
#include <algorithm>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace fg {

struct G3 {
	int nx, ny, nz;
	float h, h1, h2;
};
__device__ __forceinline__ int ix3(const G3& g, int i, int j, int k) { return i + g.nx * j + g.ny * g.nx * k; }
__device__ __forceinline__ float fmin_ref(float a, float b) { return a < b ? a : b; }   // the header's own min/max templates (:8-16)
__device__ __forceinline__ float fmax_ref(float a, float b) { return a > b ? a : b; }

// one cell of solveIncompressibility, fluid3d.h:147-180
__device__ __forceinline__ void project_cell(const G3& g, int i, int j, int k, float* __restrict__ u, float* __restrict__ v,
                                             float* __restrict__ w, float* __restrict__ pr, const unsigned char* __restrict__ solid, float cp) {
	const int c = ix3(g, i, j, k), sx = 1, sy = g.nx, sz = g.nx * g.ny;
	if (solid[c]) return;
	const int sx0 = !solid[c - sx], sx1 = !solid[c + sx], sy0 = !solid[c - sy], sy1 = !solid[c + sy], sz0 = !solid[c - sz], sz1 = !solid[c + sz];
	const int s = sx0 + sx1 + sy0 + sy1 + sz0 + sz1;
	if (s == 0) return;
	// the producers of these values are other CTAs of the same launch (wavefront): L2-coherent accesses
	const float u0 = __ldcg(u + c), u1 = __ldcg(u + c + sx), v0 = __ldcg(v + c), v1 = __ldcg(v + c + sy), w0 = __ldcg(w + c), w1 = __ldcg(w + c + sz);
	const float div = ((((u1 - u0) + v1) - v0) + w1) - w0;
	float p = -div / (float)s;
	p *= 1.9f;
	__stcg(pr + c, __ldcg(pr + c) + cp * p);
	__stcg(u + c, u0 - (float)sx0 * p); __stcg(u + c + sx, u1 + (float)sx1 * p);
	__stcg(v + c, v0 - (float)sy0 * p); __stcg(v + c + sy, v1 + (float)sy1 * p);
	__stcg(w + c, w0 - (float)sz0 * p); __stcg(w + c + sz, w1 + (float)sz1 * p);
}

// Lexicographic order (i outer, j, k inner) as a pipelined wavefront: a cell shares faces only with its six face
// neighbours, so all cells with i+j+k = const are independent and iteration n+1 may trail iteration n by two planes --
// exactly the reference's updates, in an order that respects every dependency (the 3D form of sor.cu's wavefront).
__global__ void __launch_bounds__(256) k3_project_wavefront(G3 g, int iters, float cp, float* u, float* v, float* w, float* pr,
                                                            const unsigned char* solid) {
	cg::grid_group grid = cg::this_grid();
	const int dmin = 3, dmax = (g.nx - 2) + (g.ny - 2) + (g.nz - 2);
	const int steps = (dmax - dmin + 1) + 2 * (iters - 1);
	const int ni = g.nx - 2, nj = g.ny - 2;
	for (int t = 0; t < steps; t++) {
		for (int it = blockIdx.y; it < iters; it += gridDim.y) {
			const int d = dmin + t - 2 * it;
			if (d < dmin || d > dmax) continue;
			// cells (i, j, k = d-i-j) of the plane, enumerated over (i, j)
			for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ni * nj; e += gridDim.x * blockDim.x) {
				const int i = 1 + e % ni, j = 1 + e / ni, k = d - i - j;
				if (k >= 1 && k <= g.nz - 2) project_cell(g, i, j, k, u, v, w, pr, solid, cp);
			}
		}
		grid.sync();
	}
}
// red-black order (parity of i+j+k), one colour per launch
__global__ void __launch_bounds__(256) k3_project_rb(G3 g, int colour, float cp, float* u, float* v, float* w, float* pr, const unsigned char* solid) {
	const size_t n = (size_t)(g.nx - 2) * (g.ny - 2) * (g.nz - 2);
	for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
		const int i = 1 + (int)(e % (g.nx - 2)), j = 1 + (int)((e / (g.nx - 2)) % (g.ny - 2)), k = 1 + (int)(e / ((size_t)(g.nx - 2) * (g.ny - 2)));
		if (((i + j + k) & 1) == colour) project_cell(g, i, j, k, u, v, w, pr, solid, cp);
	}
}

// extrapolate, fluid3d.h:186-210, one launch per border loop (Q21)
__global__ void __launch_bounds__(256) k3_extrapolate(G3 g, int phase, float* __restrict__ u, float* __restrict__ v, float* __restrict__ w) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (phase == 0) {          // (i, j): u, v at k = 0 and k = nz-1
		if (e >= g.nx * g.ny) return;
		const int i = e % g.nx, j = e / g.nx;
		u[ix3(g, i, j, 0)] = u[ix3(g, i, j, 1)]; u[ix3(g, i, j, g.nz - 1)] = u[ix3(g, i, j, g.nz - 2)];
		v[ix3(g, i, j, 0)] = v[ix3(g, i, j, 1)]; v[ix3(g, i, j, g.nz - 1)] = v[ix3(g, i, j, g.nz - 2)];
	} else if (phase == 1) {   // (j, k): v, w at i = 0 and i = nx-1
		if (e >= g.ny * g.nz) return;
		const int j = e % g.ny, k = e / g.ny;
		v[ix3(g, 0, j, k)] = v[ix3(g, 1, j, k)]; v[ix3(g, g.nx - 1, j, k)] = v[ix3(g, g.nx - 2, j, k)];
		w[ix3(g, 0, j, k)] = w[ix3(g, 1, j, k)]; w[ix3(g, g.nx - 1, j, k)] = w[ix3(g, g.nx - 2, j, k)];
	} else {                   // (k, i): w, u at j = 0 and j = ny-1
		if (e >= g.nz * g.nx) return;
		const int i = e % g.nx, k = e / g.nx;
		w[ix3(g, i, 0, k)] = w[ix3(g, i, 1, k)]; w[ix3(g, i, g.ny - 1, k)] = w[ix3(g, i, g.ny - 2, k)];
		u[ix3(g, i, 0, k)] = u[ix3(g, i, 1, k)]; u[ix3(g, i, g.ny - 1, k)] = u[ix3(g, i, g.ny - 2, k)];
	}
}

// sampleField, fluid3d.h:213-264 (Q20)
__device__ __forceinline__ float sample3(const G3& g, const float* __restrict__ f, float dx, float dy, float dz, float x, float y, float z) {
	x = fmax_ref(g.h, fmin_ref(x, (float)g.nx * g.h));
	y = fmax_ref(g.h, fmin_ref(y, (float)g.ny * g.h));
	z = fmax_ref(g.h, fmin_ref(z, (float)g.nz * g.h));
	const int x0 = min((int)(g.h1 * (x - dx)), g.nx - 1), y0 = min((int)(g.h1 * (y - dy)), g.ny - 1), z0 = min((int)(g.h1 * (z - dz)), g.nz - 1);
	const int x1 = min(x0 + 1, g.nx - 1), y1 = min(y0 + 1, g.ny - 1), z1 = min(z0 + 1, g.nz - 1);
	const float tx = g.h1 * ((x - dx) - g.h * (float)x0), ty = g.h1 * ((y - dy) - g.h * (float)y0), tz = g.h1 * ((z - dz) - g.h * (float)z0);
	const float sx = 1 - tx, sy = 1 - ty, sz = 1 - tz;
	return sx * sy * sz * f[ix3(g, x0, y0, z0)] + sx * sy * tz * f[ix3(g, x0, y0, z1)] + sx * ty * sz * f[ix3(g, x0, y1, z0)] +
	       sx * ty * tz * f[ix3(g, x0, y1, z1)] + tx * sy * sz * f[ix3(g, x1, y0, z0)] + tx * sy * tz * f[ix3(g, x1, y0, z1)] +
	       tx * ty * sz * f[ix3(g, x1, y1, z0)] + tx * ty * tz * f[ix3(g, x1, y1, z1)];
}
__device__ __forceinline__ float avg_u(const G3& g, const float* __restrict__ u, int i, int j, int k) {   // :266-277
	return (u[ix3(g, i, j, k)] + u[ix3(g, i, j, k - 1)] + u[ix3(g, i, j - 1, k)] + u[ix3(g, i, j - 1, k - 1)] + u[ix3(g, i + 1, j, k)] +
	        u[ix3(g, i + 1, j, k - 1)] + u[ix3(g, i + 1, j - 1, k)] + u[ix3(g, i + 1, j - 1, k - 1)]) / 8;
}
__device__ __forceinline__ float avg_v(const G3& g, const float* __restrict__ v, int i, int j, int k) {   // :279-290
	return (v[ix3(g, i, j, k)] + v[ix3(g, i, j, k - 1)] + v[ix3(g, i - 1, j, k)] + v[ix3(g, i - 1, j, k - 1)] + v[ix3(g, i, j + 1, k)] +
	        v[ix3(g, i, j + 1, k - 1)] + v[ix3(g, i - 1, j + 1, k)] + v[ix3(g, i - 1, j + 1, k - 1)]) / 8;
}
__device__ __forceinline__ float avg_w(const G3& g, const float* __restrict__ w, int i, int j, int k) {   // :292-303
	return (w[ix3(g, i, j, k)] + w[ix3(g, i, j - 1, k)] + w[ix3(g, i - 1, j, k)] + w[ix3(g, i - 1, j - 1, k)] + w[ix3(g, i, j, k + 1)] +
	        w[ix3(g, i, j - 1, k + 1)] + w[ix3(g, i - 1, j, k + 1)] + w[ix3(g, i - 1, j - 1, k + 1)]) / 8;
}

// advectVel, fluid3d.h:305-346: every cell of the new fields is written once (the old value where the reference leaves it)
__global__ void __launch_bounds__(256) k3_advect_vel(G3 g, float dt, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
                                                     const unsigned char* __restrict__ solid, float* __restrict__ nu, float* __restrict__ nv,
                                                     float* __restrict__ nw) {
	const size_t n = (size_t)g.nx * g.ny * g.nz;
	const int kmax = min(g.ny, g.nz);   // Q19
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
		const int i = (int)(c % g.nx), j = (int)((c / g.nx) % g.ny), k = (int)(c / ((size_t)g.nx * g.ny));
		float ou = u[c], ov = v[c], ow = w[c];
		if (i >= 1 && j >= 1 && k >= 1 && k < kmax && !solid[c]) {
			if (!solid[ix3(g, i - 1, j, k)] && j < g.ny - 1 && k < g.nz - 1) {
				const float x = g.h * (float)i, y = g.h2 + g.h * (float)j, z = g.h2 + g.h * (float)k;
				const float u0 = u[c], v0 = avg_v(g, v, i, j, k), w0 = avg_w(g, w, i, j, k);
				ou = sample3(g, u, 0.f, g.h2, 0.f, x - dt * u0, y - dt * v0, z - dt * w0);
			}
			if (!solid[ix3(g, i, j - 1, k)] && k < g.nz - 1 && i < g.nx - 1) {
				const float x = g.h2 + g.h * (float)i, y = g.h * (float)j, z = g.h2 + g.h * (float)k;
				const float u0 = avg_u(g, u, i, j, k), v0 = v[c], w0 = avg_w(g, w, i, j, k);
				ov = sample3(g, v, g.h2, 0.f, 0.f, x - dt * u0, y - dt * v0, z - dt * w0);
			}
			if (!solid[ix3(g, i, j, k - 1)] && i < g.nx - 1 && j < g.ny - 1) {
				const float x = g.h2 + g.h * (float)i, y = g.h2 + g.h * (float)j, z = g.h * (float)k;
				const float u0 = avg_u(g, u, i, j, k), v0 = avg_v(g, v, i, j, k), w0 = w[c];
				ow = sample3(g, w, 0.f, 0.f, g.h2, x - dt * u0, y - dt * v0, z - dt * w0);
			}
		}
		nu[c] = ou; nv[c] = ov; nw[c] = ow;
	}
}
// advectSmoke, fluid3d.h:348-367
__global__ void __launch_bounds__(256) k3_advect_smoke(G3 g, float dt, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
                                                       const float* __restrict__ m, const unsigned char* __restrict__ solid, float* __restrict__ nm) {
	const size_t n = (size_t)g.nx * g.ny * g.nz;
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
		const int i = (int)(c % g.nx), j = (int)((c / g.nx) % g.ny), k = (int)(c / ((size_t)g.nx * g.ny));
		float om = m[c];
		if (i >= 1 && i < g.nx - 1 && j >= 1 && j < g.ny - 1 && k >= 1 && k < g.nz - 1 && !solid[c]) {
			const float u0 = (u[c] + u[c + 1]) / 2, v0 = (v[c] + v[c + g.nx]) / 2, w0 = (w[c] + w[c + g.nx * g.ny]) / 2;
			const float x = g.h2 + g.h * (float)i - dt * u0, y = g.h2 + g.h * (float)j - dt * v0, z = g.h2 + g.h * (float)k - dt * w0;
			om = sample3(g, m, g.h2, g.h2, 0.f, x, y, z);
		}
		nm[c] = om;
	}
}
__global__ void k3_fill(size_t n, float* a, float v) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void k3_sample(G3 g, const float* f, float dx, float dy, float dz, const float* xs, const float* ys, const float* zs, float* out, size_t n) {
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = sample3(g, f, dx, dy, dz, xs[i], ys[i], zs[i]);
}

}  // namespace fg

using namespace fg;

struct Fluid3DSim {
	Fluid3DDims d{};
	int device = 0;
	cudaStream_t st = nullptr;
	float *u = nullptr, *v = nullptr, *w = nullptr, *new_u = nullptr, *new_v = nullptr, *new_w = nullptr, *pressure = nullptr, *m = nullptr, *new_m = nullptr;
	unsigned char* solid = nullptr;
	LaunchCounter lc;
};

namespace {
struct DG3 {
	int prev = -1;
	explicit DG3(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
	~DG3() { if (prev >= 0) cudaSetDevice(prev); }
};
G3 g3(const Fluid3DSim* f) { return G3{f->d.num_x, f->d.num_y, f->d.num_z, f->d.h, 1 / f->d.h, f->d.h / 2}; }
float** field3(Fluid3DSim* f, int field) {
	switch (field) {
		case FLUID3D_U: return &f->u; case FLUID3D_V: return &f->v; case FLUID3D_W: return &f->w;
		case FLUID3D_NEW_U: return &f->new_u; case FLUID3D_NEW_V: return &f->new_v; case FLUID3D_NEW_W: return &f->new_w;
		case FLUID3D_PRESSURE: return &f->pressure; case FLUID3D_M: return &f->m; case FLUID3D_NEW_M: return &f->new_m;
	}
	return nullptr;
}
}  // namespace

extern "C" {

int fluid3d_create(int x, int y, int z, float density, float h, int device, Fluid3DSim** out) {
	FG_CHECK(out && x >= 1 && y >= 1 && z >= 1 && h > 0, FLIPGPU_EINVAL, "fluid3d_create: bad argument");
	FG_CHECK((long long)(x + 2) * (y + 2) * (z + 2) < (1ll << 31), FLIPGPU_EINVAL, "fluid3d_create: grid too large for int32 indices");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		fg::set_error("fluid3d_create: no CUDA device (this library has no CPU fallback)");
		return FLIPGPU_ENODEV;
	}
	if (device < 0) FG_CUDA(cudaGetDevice(&device));
	FG_CHECK(device < ndev, FLIPGPU_EINVAL, "fluid3d_create: device %d of %d", device, ndev);
	Fluid3DSim* f = new Fluid3DSim;
	f->device = device;
	DG3 guard(device);
	f->d.num_x = 2 + x; f->d.num_y = 2 + y; f->d.num_z = 2 + z; f->d.num_cells = f->d.num_x * f->d.num_y * f->d.num_z;   // fluid3d.h:42-43
	f->d.density = density; f->d.h = h;
	FG_CUDA(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
	const size_t C = f->d.num_cells;
	float** fs[] = {&f->u, &f->v, &f->w, &f->new_u, &f->new_v, &f->new_w, &f->pressure, &f->m, &f->new_m};
	for (float** p : fs) {
		FG_CUDA(cudaMalloc(p, C * 4));
		FG_CUDA(cudaMemsetAsync(*p, 0, C * 4, f->st));
	}
	FG_CUDA(cudaMalloc(&f->solid, C));
	FG_CUDA(cudaMemsetAsync(f->solid, 0, C, f->st));
	k3_fill<<<wave_grid(C, 256), 256, 0, f->st>>>(C, f->m, 1.f);   // :64-68
	f->lc.n++;
	FG_CUDA(cudaStreamSynchronize(f->st));
	*out = f;
	return FLIPGPU_OK;
}
int fluid3d_destroy(Fluid3DSim* f) {
	if (!f) return FLIPGPU_OK;
	DG3 guard(f->device);
	cudaStreamSynchronize(f->st);
	void* ptrs[] = {f->u, f->v, f->w, f->new_u, f->new_v, f->new_w, f->pressure, f->m, f->new_m, f->solid};
	for (void* p : ptrs) if (p) cudaFree(p);
	cudaStreamDestroy(f->st);
	delete f;
	return FLIPGPU_OK;
}
int fluid3d_get_dims(Fluid3DSim* f, Fluid3DDims* dims) {
	FG_CHECK(f && dims, FLIPGPU_EINVAL, "null argument");
	*dims = f->d;
	return FLIPGPU_OK;
}
static int copy3(Fluid3DSim* f, int field, void* host, size_t count, bool up) {
	FG_CHECK(f && host, FLIPGPU_EINVAL, "fluid3d transfer: null argument");
	FG_CHECK(field >= 0 && field < FLUID3D_FIELD_COUNT, FLIPGPU_EINVAL, "fluid3d transfer: unknown field %d", field);
	FG_CHECK(count == (size_t)f->d.num_cells, FLIPGPU_EINVAL, "fluid3d transfer: expected %d elements, got %zu", f->d.num_cells, count);
	DG3 guard(f->device);
	void* dev = field == FLUID3D_SOLID ? (void*)f->solid : (void*)*field3(f, field);
	const size_t bytes = field == FLUID3D_SOLID ? count : count * 4;
	FG_CUDA(cudaMemcpyAsync(up ? dev : host, up ? host : dev, bytes, up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, f->st));
	FG_CUDA(cudaStreamSynchronize(f->st));
	return FLIPGPU_OK;
}
int fluid3d_upload(Fluid3DSim* f, int field, const void* host, size_t count) { return copy3(f, field, const_cast<void*>(host), count, true); }
int fluid3d_readback(Fluid3DSim* f, int field, void* host, size_t count) { return copy3(f, field, host, count, false); }

int fluid3d_project(Fluid3DSim* f, int iters, float dt, int order) {
	FG_CHECK(f && iters >= 0 && dt > 0, FLIPGPU_EINVAL, "fluid3d_project: bad argument");
	FG_CHECK(order == FLIP_SOLVER_LEX_WAVEFRONT || order == FLIP_SOLVER_RED_BLACK, FLIPGPU_EINVAL, "fluid3d_project: order %d (lexicographic = 0 or red-black = 1)", order);
	DG3 guard(f->device);
	G3 g = g3(f);
	const float cp = f->d.density * f->d.h / dt;   // :137
	FG_CUDA(cudaMemsetAsync(f->pressure, 0, (size_t)f->d.num_cells * 4, f->st));   // :138-140
	if (iters == 0 || g.nx < 3 || g.ny < 3 || g.nz < 3) return FLIPGPU_OK;
	if (order == FLIP_SOLVER_RED_BLACK) {
		const size_t n = (size_t)(g.nx - 2) * (g.ny - 2) * (g.nz - 2);
		for (int it = 0; it < iters; it++)
			for (int colour = 1; colour >= 0; colour--) {   // (1,1,1) has odd parity: start with the colour of the first cell
				k3_project_rb<<<wave_grid(n, 256), 256, 0, f->st>>>(g, colour, cp, f->u, f->v, f->w, f->pressure, f->solid);
				f->lc.n++;
			}
		FG_CUDA(cudaGetLastError());
		return FLIPGPU_OK;
	}
	int per_sm = 0;
	FG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k3_project_wavefront, 256, 0));
	const int total = std::max(1, per_sm) * num_sms();
	// blockIdx.y = iterations in flight (a plane holds at most (nx-2)(ny-2) cells), blockIdx.x = cells of a plane
	const int by = std::max(1, std::min(iters, std::min(total, 16)));
	const int bx = std::max(1, std::min(total / by, (int)cdiv((size_t)(g.nx - 2) * (g.ny - 2), 256)));
	float *u = f->u, *v = f->v, *w = f->w, *pr = f->pressure;
	const unsigned char* solid = f->solid;
	void* args[] = {&g, &iters, (void*)&cp, &u, &v, &w, &pr, &solid};
	FG_CUDA(cudaLaunchCooperativeKernel((const void*)k3_project_wavefront, dim3(bx, by), dim3(256), args, 0, f->st));
	f->lc.n++;
	return FLIPGPU_OK;
}
int fluid3d_extrapolate(Fluid3DSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DG3 guard(f->device);
	G3 g = g3(f);
	const int n[3] = {g.nx * g.ny, g.ny * g.nz, g.nz * g.nx};
	for (int ph = 0; ph < 3; ph++) {
		k3_extrapolate<<<cdiv(n[ph], 256), 256, 0, f->st>>>(g, ph, f->u, f->v, f->w);
		f->lc.n++;
	}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}
int fluid3d_advect_vel(Fluid3DSim* f, float dt) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DG3 guard(f->device);
	k3_advect_vel<<<wave_grid(f->d.num_cells, 256), 256, 0, f->st>>>(g3(f), dt, f->u, f->v, f->w, f->solid, f->new_u, f->new_v, f->new_w);
	f->lc.n++;
	FG_CUDA(cudaGetLastError());
	std::swap(f->u, f->new_u); std::swap(f->v, f->new_v); std::swap(f->w, f->new_w);   // ping-pong instead of the reference's two memcpys
	return FLIPGPU_OK;
}
int fluid3d_advect_smoke(Fluid3DSim* f, float dt) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DG3 guard(f->device);
	k3_advect_smoke<<<wave_grid(f->d.num_cells, 256), 256, 0, f->st>>>(g3(f), dt, f->u, f->v, f->w, f->m, f->solid, f->new_m);
	f->lc.n++;
	FG_CUDA(cudaGetLastError());
	std::swap(f->m, f->new_m);
	return FLIPGPU_OK;
}
int fluid3d_step(Fluid3DSim* f, int iters, float dt, int order, int n_steps) {
	for (int k = 0; k < n_steps; k++) {   // the 2D caller sequence (eulerian_fluid/src/fluid_ui.h:107-111) carried over
		FG_TRY(fluid3d_project(f, iters, dt, order));
		FG_TRY(fluid3d_extrapolate(f));
		FG_TRY(fluid3d_advect_vel(f, dt));
		FG_TRY(fluid3d_advect_smoke(f, dt));
	}
	return FLIPGPU_OK;
}
int fluid3d_sample_field(Fluid3DSim* f, int field, const float* xs, const float* ys, const float* zs, float* out, size_t n) {
	FG_CHECK(f && xs && ys && zs && out, FLIPGPU_EINVAL, "fluid3d_sample_field: null argument");
	FG_CHECK(field >= 0 && field <= 3, FLIPGPU_EINVAL, "fluid3d_sample_field: field %d (0 U, 1 V, 2 W, 3 S)", field);
	if (n == 0) return FLIPGPU_OK;
	DG3 guard(f->device);
	G3 g = g3(f);
	float* d = nullptr;
	FG_CUDA(cudaMalloc(&d, 4 * n * 4));
	cudaMemcpyAsync(d, xs, n * 4, cudaMemcpyHostToDevice, f->st);
	cudaMemcpyAsync(d + n, ys, n * 4, cudaMemcpyHostToDevice, f->st);
	cudaMemcpyAsync(d + 2 * n, zs, n * 4, cudaMemcpyHostToDevice, f->st);
	const float* src = field == 0 ? f->u : field == 1 ? f->v : field == 2 ? f->w : f->m;
	const float dx = (field == 1 || field == 3) ? g.h2 : 0.f, dy = (field == 0 || field == 3) ? g.h2 : 0.f, dz = field == 2 ? g.h2 : 0.f;   // :228-233
	k3_sample<<<wave_grid(n, 256), 256, 0, f->st>>>(g, src, dx, dy, dz, d, d + n, d + 2 * n, d + 3 * n, n);
	f->lc.n++;
	cudaError_t err = cudaMemcpyAsync(out, d + 3 * n, n * 4, cudaMemcpyDeviceToHost, f->st);
	if (err == cudaSuccess) err = cudaStreamSynchronize(f->st);
	cudaFree(d);
	FG_CUDA(err);
	return FLIPGPU_OK;
}
int fluid3d_synchronize(Fluid3DSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DG3 guard(f->device);
	FG_CUDA(cudaStreamSynchronize(f->st));
	return FLIPGPU_OK;
}
int fluid3d_get_launch_count(Fluid3DSim* f, long long* launches) {
	FG_CHECK(f && launches, FLIPGPU_EINVAL, "null argument");
	*launches = f->lc.n;
	return FLIPGPU_OK;
}

}  // extern "C"
