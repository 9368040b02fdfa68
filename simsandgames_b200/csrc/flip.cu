// B200-native FLIP step behind the C ABI of include/flipgpu.h.
// Replaces struct FlipFluid (reference flip_fluid/src/flip_fluid.h:13-582); every kernel cites the
// reference lines whose arithmetic it reproduces.  Compiled with -fmad=false: fp32 products and sums
// stay unfused and in the reference's association order, so per-particle phases are bit-exact against
// the serial CPU step and only the ORDER of some reductions differs in the production path (P2G adds,
// separation sweep, rest-density sum) -- flip_set_exact_order() switches those to the reference's order too.
//
// One simulate() = integrate+histogram -> running sum, fill, per-cell fix-up, physical reorder (F1,F2)
//   -> 2 Jacobi separation sweeps (F3) -> collide + mark + base-cell histogram (F4,F5)
//   -> per-cell slot lists -> gather-form P2G + density + rest latch (F6,F7) -> SOR solve (F9, sor*.cu)
//   -> G2P (F8) [-> colours (F10)].   The slab-sharded step (shard_step) wraps the same kernels.
//
// Data layout in HBM:
//   grid   : u v prev_u prev_v p s density  float[C], cell_type int[C], x-fastest like fIX (flip_fluid.h:147)
//   particles: pos/vel float2[P] x2 (ping-pong), physically sorted by spatial-hash cell (pIX order,
//            descending caller id inside a cell -- exactly cell_particle_ids order, flip_fluid.h:196-197);
//            orig[] = caller index of each slot == cell_particle_ids.
//   hash   : num_cell_particles int[Hc], first_cell_particle int[Hc+1]  (bit-exact vs flip_fluid.h:169-198)
//   P2G    : gcount int[C], gfirst int[C+1] + slot lists per stencil base cell (scratch reused from the sort)
#include <algorithm>
#include <cmath>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"
#include "nccl_shim.cuh"

namespace fg {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

constexpr int kThreads = 256;
constexpr int kScanThreads = 512, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

struct Geo {  // constants every kernel needs, passed by value
	int nx, ny, pnx, pny, np;
	float h, inv_h, h2, r, p_inv;
	float xmax_c, ymax_c;  // h*(nx-1), h*(ny-1): clamp bounds of the bilinear stencil (flip_fluid.h:301-302)
	// Row windows.  One GPU: the whole grid.  Slab-sharded: the rows this rank keeps current (own rows + halo);
	// particles that stray outside (invalid outer ghosts) are folded into the edge row so that every index stays
	// inside the rows whose counters are cleared and scanned.
	int gy0, gy1;          // grid rows [gy0, gy1)
	int hy0, hy1;          // spatial-hash rows [hy0, hy1)
	int ny0, ny1;          // node rows the P2G gather produces
};

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
__device__ __forceinline__ float clampf(float a, float lo, float hi) { return a < lo ? lo : (hi < a ? hi : a); }

// spatial-hash cell, flip_fluid.h:174-176
__device__ __forceinline__ int hash_cell(const Geo& g, float x, float y, int* cx = nullptr, int* cy = nullptr) {
	int xi = clampi((int)(x * g.p_inv), 0, g.pnx - 1);
	int yi = clampi((int)(y * g.p_inv), g.hy0, g.hy1 - 1);
	if (cx) { *cx = xi; *cy = yi; }
	return xi + g.pnx * yi;
}

// bilinear stencil of flip_fluid.h:301-313 == :374-396 (Q2: same cell-centred offset for u, v and density)
struct Stencil {
	int x0, x1, y0, y1;
	float w0, w1, w2, w3;  // d0..d3 of flip_fluid.h:388-391 for nodes (x0,y0) (x1,y0) (x1,y1) (x0,y1)
};
__device__ __forceinline__ Stencil make_stencil(const Geo& g, float x, float y) {
	x = clampf(x, g.h, g.xmax_c);
	y = clampf(y, g.h, g.ymax_c);
	Stencil st;
	float xs = x - g.h2, ys = y - g.h2;
	st.x0 = (int)(g.inv_h * xs);
	float tx = g.inv_h * (xs - g.h * (float)st.x0);
	st.x1 = min(1 + st.x0, g.nx - 2);
	st.y0 = (int)(g.inv_h * ys);
	float ty = g.inv_h * (ys - g.h * (float)st.y0);
	st.y1 = min(1 + st.y0, g.ny - 2);
	float sx = 1.f - tx, sy = 1.f - ty;
	st.w0 = sx * sy; st.w1 = tx * sy; st.w2 = tx * ty; st.w3 = sx * ty;
	return st;
}

__device__ __forceinline__ int warp_max(int v) {
	for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// ------------------------------------------------------------------ F1 (+ first half of F2)
// integrateParticles (flip_fluid.h:156-162) fused with the key + histogram of the counting sort (:169-177).
// STORE = false (whole steps only): the integrated state is not written back -- only the sort key leaves the kernel -- and
// k_reorder<.., true> redoes the same three operations on the values it gathers: 16 of the 36 bytes per particle less.
__device__ __forceinline__ void integrate_one(float2& p, float2& v, float dt, float gravity) {
	v.y += gravity * dt;
	p.x += v.x * dt;
	p.y += v.y * dt;
}
template <bool INTEGRATE, bool BIN, bool STORE = true>
__global__ void __launch_bounds__(kThreads) k_integrate_bin(Geo g, float2* __restrict__ pos, float2* __restrict__ vel,
                                                            float dt, float gravity, int* __restrict__ keys,
                                                            int* __restrict__ count) {
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < g.np; k += gridDim.x * blockDim.x) {
		float2 p = pos[k];
		if (INTEGRATE) {
			float2 v = vel[k];
			integrate_one(p, v, dt, gravity);
			if (STORE) { vel[k] = v; pos[k] = p; }
		}
		if (BIN) {
			int key = hash_cell(g, p.x, p.y);
			keys[k] = key;
			atomicAdd(count + key, 1);  // integer histogram: the result is order-independent
		}
	}
}

// ------------------------------------------------------------------ F2: running sum (flip_fluid.h:180-186)
// Three passes over the Hc counts: tile sums, scan of tile sums, tile rescan.  first[c] leaves this
// stage INCLUSIVE (as :183); the fill's decrements turn it into the exclusive start (as :196).
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const int* __restrict__ count, int n, int* __restrict__ tile_sum) {
	size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
	int s = 0;
	const bool vec = (reinterpret_cast<size_t>(count) & 15) == 0;  // window offsets need not be 16-byte aligned
	if (vec && base + kScanItems <= (size_t)n) {
		int4 a = *reinterpret_cast<const int4*>(count + base), b = *reinterpret_cast<const int4*>(count + base + 4);
		s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
	} else {
		for (int i = 0; i < kScanItems; i++) if (base + i < (size_t)n) s += count[base + i];
	}
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	__shared__ int ws[kScanThreads / 32];
	if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int i = 0; i < kScanThreads / 32; i++) t += ws[i];
		tile_sum[blockIdx.x] = t;
	}
}

__global__ void __launch_bounds__(1024) k_scan_tile_offsets(int* __restrict__ tile_sum, int n_tiles) {
	// single CTA: exclusive scan of the tile sums in place, 1024 at a time with a running carry
	__shared__ int ws[32];
	__shared__ int carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (int base = 0; base < n_tiles; base += 1024) {
		int i = base + threadIdx.x;
		int v = i < n_tiles ? tile_sum[i] : 0;
		int incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, incl, o);
			if ((threadIdx.x & 31) >= o) incl += t;
		}
		if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
		__syncthreads();
		if (threadIdx.x < 32) {
			int w = ws[threadIdx.x], wi = w;
			for (int o = 1; o < 32; o <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, wi, o);
				if (threadIdx.x >= o) wi += t;
			}
			ws[threadIdx.x] = wi - w;  // exclusive warp offsets
		}
		__syncthreads();
		int carry = carry_s;
		int excl = carry + ws[threadIdx.x >> 5] + incl - v;
		if (i < n_tiles) tile_sum[i] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) carry_s = excl + v;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const int* __restrict__ count, int n,
                                                             const int* __restrict__ tile_off, int* __restrict__ first) {
	size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
	int v[kScanItems];
	const bool vec = ((reinterpret_cast<size_t>(count) | reinterpret_cast<size_t>(first)) & 15) == 0;
	if (vec && base + kScanItems <= (size_t)n) {
		int4 a = *reinterpret_cast<const int4*>(count + base), b = *reinterpret_cast<const int4*>(count + base + 4);
		v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
	} else {
		for (int i = 0; i < kScanItems; i++) v[i] = (base + i < (size_t)n) ? count[base + i] : 0;
	}
	for (int i = 1; i < kScanItems; i++) v[i] += v[i - 1];
	int tot = v[kScanItems - 1], incl = tot;
	for (int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, incl, o);
		if ((threadIdx.x & 31) >= o) incl += t;
	}
	__shared__ int ws[kScanThreads / 32];
	if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
	__syncthreads();
	if (threadIdx.x < 32) {
		int w = threadIdx.x < kScanThreads / 32 ? ws[threadIdx.x] : 0, wi = w;
		for (int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, wi, o);
			if (threadIdx.x >= o) wi += t;
		}
		if (threadIdx.x < kScanThreads / 32) ws[threadIdx.x] = wi - w;
	}
	__syncthreads();
	int off = tile_off[blockIdx.x] + ws[threadIdx.x >> 5] + incl - tot;
	if (vec && base + kScanItems <= (size_t)n) {
		int4 a = make_int4(v[0] + off, v[1] + off, v[2] + off, v[3] + off);
		int4 b = make_int4(v[4] + off, v[5] + off, v[6] + off, v[7] + off);
		*reinterpret_cast<int4*>(first + base) = a;
		*reinterpret_cast<int4*>(first + base + 4) = b;
		if (base + kScanItems == (size_t)n) first[n] = v[7] + off;  // guard entry, flip_fluid.h:186
	} else {
		for (int i = 0; i < kScanItems; i++)
			if (base + i < (size_t)n) {
				first[base + i] = v[i] + off;
				if (base + i == (size_t)n - 1) first[n] = v[i] + off;
			}
	}
}


// Sort a short run held in registers with an odd-even transposition network (fully unrolled, compile-time
// indices).  Cell segments hold a handful of particles, so this replaces the read-modify-write insertion sort
// on global memory for every run of up to N elements; longer runs fall back to the in-place insertion sort.
template <int N, typename T>
__device__ __forceinline__ void sort_network_ascending(T (&v)[N]) {
#pragma unroll
	for (int r = 0; r < N; r++) {
#pragma unroll
		for (int i = (r & 1); i + 1 < N; i += 2) {
			T lo = v[i] < v[i + 1] ? v[i] : v[i + 1], hi = v[i] < v[i + 1] ? v[i + 1] : v[i];
			v[i] = lo; v[i + 1] = hi;
		}
	}
}

// ------------------------------------------------------------------ F2: fill (flip_fluid.h:189-198)
// Slots are claimed with atomicSub (any order) ...
__global__ void __launch_bounds__(kThreads) k_fill(int np, const int* __restrict__ keys, const int* __restrict__ orig,
                                                   int* __restrict__ first, int* __restrict__ src, int* __restrict__ ids) {
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
		const int key = keys[k];
		if (key < 0) continue;   // kDeadKey: a particle this rank no longer holds (sharded steps) drops out of the sort
		int slot = atomicSub(first + key, 1) - 1;
		src[slot] = k;
		ids[slot] = orig[k];
	}
}
// ... and each cell's segment is then put into the reference's order: DESCENDING caller index (Q7),
// which is what the serial decrement-on-fill of :196-197 produces.  Deterministic and bit-exact.
// Segments longer than kCrowd (particles piled up against a wall: hundreds per cell in a developed flow) are not sorted by
// their one thread -- an insertion sort of 535 entries through global memory kept the whole kernel waiting for 1.5 ms at S2 --
// but queued for k_fixup_crowded, which ranks a segment with a whole CTA.  A full queue falls back to the in-thread sort.
constexpr int kCrowd = 16, kCrowdMax = 4096, kCrowdCap = 1 << 20;
struct CrowdQueue { int* list; int* count; };
__device__ __forceinline__ bool crowd_push(const CrowdQueue& q, int c) {
	const int k = atomicAdd(q.count, 1);
	if (k >= kCrowdCap) return false;
	q.list[k] = c;
	return true;
}
__global__ void __launch_bounds__(kThreads) k_fixup_cells(int n_cells, const int* __restrict__ count,
                                                          const int* __restrict__ first, int* __restrict__ src,
                                                          int* __restrict__ ids, CrowdQueue crowd) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
		int n = count[c];
		if (n < 2) continue;
		if (n > kCrowd && crowd_push(crowd, c)) continue;
		int b = first[c];
		// (id, src) packed into one 64-bit key, negated: ascending order of the key == descending id (ids are unique)
		if (n == 2) {
			int i0 = ids[b], i1 = ids[b + 1];
			if (i0 < i1) { int s0 = src[b], s1 = src[b + 1]; ids[b] = i1; ids[b + 1] = i0; src[b] = s1; src[b + 1] = s0; }
			continue;
		}
		if (n <= 4) {
			long long v[4];
#pragma unroll
			for (int i = 0; i < 4; i++)
				v[i] = i < n ? -(((long long)ids[b + i] << 32) | (unsigned)src[b + i]) : 0x7fffffffffffffffll;
			sort_network_ascending<4>(v);
#pragma unroll
			for (int i = 0; i < 4; i++)
				if (i < n) { long long w = -v[i]; ids[b + i] = (int)(w >> 32); src[b + i] = (int)(w & 0xffffffffll); }
			continue;
		}
		if (n <= 8) {
			long long v[8];
#pragma unroll
			for (int i = 0; i < 8; i++)
				v[i] = i < n ? -(((long long)ids[b + i] << 32) | (unsigned)src[b + i]) : 0x7fffffffffffffffll;
			sort_network_ascending<8>(v);
#pragma unroll
			for (int i = 0; i < 8; i++)
				if (i < n) { long long w = -v[i]; ids[b + i] = (int)(w >> 32); src[b + i] = (int)(w & 0xffffffffll); }
			continue;
		}
		for (int i = 1; i < n; i++) {  // insertion sort, descending id (crowded cells only)
			int id = ids[b + i], sv = src[b + i];
			int j = i - 1;
			while (j >= 0 && ids[b + j] < id) {
				ids[b + j + 1] = ids[b + j];
				src[b + j + 1] = src[b + j];
				j--;
			}
			ids[b + j + 1] = id;
			src[b + j + 1] = sv;
		}
	}
}

// The queued segments, one CTA each: rank sort in shared memory (ids and slot numbers are unique, so the rank of an entry is the
// number of entries that precede it in the order).  CELLS: (id, src) pairs by DESCENDING id (k_fixup_cells); else slot numbers
// ascending (k_fixup_grid).  Same result as the in-thread sorts.  Longer than kCrowdMax: thread 0 sorts by insertion.
template <bool CELLS>
__global__ void __launch_bounds__(256) k_fixup_crowded(CrowdQueue crowd, const int* __restrict__ count, const int* __restrict__ first,
                                                       int* __restrict__ a /* src (CELLS) or the slot list */, int* __restrict__ ids) {
	__shared__ int s_key[kCrowdMax];
	__shared__ int s_val[CELLS ? kCrowdMax : 1];
	const int total = min(*crowd.count, kCrowdCap);
	for (int i = blockIdx.x; i < total; i += gridDim.x) {
		const int c = crowd.list[i], n = count[c], b = first[c];
		if (n > kCrowdMax) {
			if (threadIdx.x == 0) {
				for (int k = 1; k < n; k++) {
					const int key = CELLS ? ids[b + k] : a[b + k], val = CELLS ? a[b + k] : 0;
					int j = k - 1;
					while (j >= 0 && (CELLS ? ids[b + j] < key : a[b + j] > key)) {
						if (CELLS) { ids[b + j + 1] = ids[b + j]; }
						a[b + j + 1] = a[b + j];
						j--;
					}
					if (CELLS) { ids[b + j + 1] = key; a[b + j + 1] = val; } else a[b + j + 1] = key;
				}
			}
			continue;
		}
		for (int t = threadIdx.x; t < n; t += blockDim.x) {
			s_key[t] = CELLS ? ids[b + t] : a[b + t];
			if (CELLS) s_val[t] = a[b + t];
		}
		__syncthreads();
		for (int t = threadIdx.x; t < n; t += blockDim.x) {
			const int key = s_key[t];
			int r = 0;
			for (int j = 0; j < n; j++) r += CELLS ? (s_key[j] > key) : (s_key[j] < key);
			if (CELLS) { ids[b + r] = key; a[b + r] = s_val[t]; } else a[b + r] = key;
		}
		__syncthreads();
	}
}

// physical reorder into hash-cell order
template <bool COLORS, bool INTEGRATE = false>
__global__ void __launch_bounds__(kThreads) k_reorder(Geo g, const int* __restrict__ src, const float2* __restrict__ pos_in,
                                                      const float2* __restrict__ vel_in, const float* __restrict__ col_in,
                                                      float2* __restrict__ pos_out, float2* __restrict__ vel_out,
                                                      float* __restrict__ col_out, float dt, float gravity) {
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		int k = src[t];
		float2 p = pos_in[k], v = vel_in[k];
		if (INTEGRATE) integrate_one(p, v, dt, gravity);   // F1 of this step, deferred by k_integrate_bin<true, true, false>
		pos_out[t] = p;
		vel_out[t] = v;
		if (COLORS) {
			col_out[3 * (size_t)t] = col_in[3 * (size_t)k];
			col_out[3 * (size_t)t + 1] = col_in[3 * (size_t)k + 1];
			col_out[3 * (size_t)t + 2] = col_in[3 * (size_t)k + 2];
		}
	}
}

// slots back to caller order (used by upload/readback of particle arrays)
template <typename T, int W>
__global__ void __launch_bounds__(kThreads) k_scatter_to_caller(int np, const int* __restrict__ orig,
                                                                const T* __restrict__ in, T* __restrict__ out) {
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < np; t += gridDim.x * blockDim.x) {
		size_t o = (size_t)orig[t];
		for (int w = 0; w < W; w++) out[W * o + w] = in[W * (size_t)t + w];
	}
}
__global__ void k_iota(int n, int* a) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = i;
}
__global__ void k_init_colors(int n, float* c) {  // flip_fluid.h:94-96
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		c[3 * (size_t)i] = 0.f; c[3 * (size_t)i + 1] = 0.f; c[3 * (size_t)i + 2] = 1.f;
	}
}

// ------------------------------------------------------------------ F4 / F5 per-particle pieces (used by k_collide_mark and the fused last separation sweep)
struct CollideArgs {
	float ox, oy, ovx, ovy, md2;          // obstacle centre, velocity, (r + orad)^2
	float min_x, max_x, min_y, max_y;     // tank walls for particle centres, flip_fluid.h:256-259
};
__device__ __forceinline__ CollideArgs collide_args(const Geo& g, float ox, float oy, float ovx, float ovy, float orad) {
	CollideArgs a;
	const float md = g.r + orad;
	a.ox = ox; a.oy = oy; a.ovx = ovx; a.ovy = ovy; a.md2 = md * md;
	a.min_x = g.h + g.r; a.max_x = g.h * (float)(g.nx - 1) - g.r;
	a.min_y = g.h + g.r; a.max_y = g.h * (float)(g.ny - 1) - g.r;
	return a;
}
// handleParticleCollisions for one particle (flip_fluid.h:262-288); returns true if the velocity changed
__device__ __forceinline__ bool collide_one(const CollideArgs& a, float2& p, float2& v) {
	bool hit = false;
	float dx = p.x - a.ox, dy = p.y - a.oy;
	if (dx * dx + dy * dy < a.md2) { v.x = a.ovx; v.y = a.ovy; hit = true; }  // Q12: velocity only, no push-out
	if (p.x < a.min_x) { p.x = a.min_x; v.x = 0.f; hit = true; }
	if (p.x > a.max_x) { p.x = a.max_x; v.x = 0.f; hit = true; }
	if (p.y < a.min_y) { p.y = a.min_y; v.y = 0.f; hit = true; }
	if (p.y > a.max_y) { p.y = a.max_y; v.y = 0.f; hit = true; }
	return hit;
}
// fluid-cell marking (flip_fluid.h:352-359) + key and histogram of the index-only sort by stencil base cell (:301-308)
__device__ __forceinline__ void mark_one(const Geo& g, float2 p, int t, int* __restrict__ cell_type, int* __restrict__ gkeys, int* __restrict__ gcount) {
	int xi = clampi((int)(g.inv_h * p.x), 0, g.nx - 1);
	int yi = clampi((int)(g.inv_h * p.y), g.gy0, g.gy1 - 1);
	int c = xi + g.nx * yi;
	if (cell_type[c] == AIR_CELL) cell_type[c] = FLUID_CELL;  // all racing writers store the same value
	float x = clampf(p.x, g.h, g.xmax_c), y = clampf(p.y, g.h, g.ymax_c);
	int x0 = clampi((int)(g.inv_h * (x - g.h2)), 0, g.nx - 1), y0 = clampi((int)(g.inv_h * (y - g.h2)), g.gy0, g.gy1 - 1);
	int key = x0 + g.nx * y0;
	gkeys[t] = key;
	atomicAdd(gcount + key, 1);  // integer histogram: order-independent
}

// ------------------------------------------------------------------ F3: separation sweep (flip_fluid.h:201-250)
// The reference sweep is a serial Gauss-Seidel over particles in caller order with an asymmetric pair
// update (Q1); no parallel schedule reproduces it.  This is the deterministic parallel schedule: one
// Jacobi sweep per launch -- every particle gathers, from the positions at the start of the sweep, the
// symmetric half-overlap push of :231-235 from each neighbour in its 3x3 hash window (window from the
// CURRENT position as :208-213, bins from before the first sweep as Q6).  Neighbours are visited
// row by row in slot order; oracle/flip_oracle.c:orcflip_separate_jacobi is the same schedule on the CPU
// and this kernel is bit-exact against it.  Statistical distance to the serial sweep: tests T3.
// FUSE: the last sweep of a step also runs the collision pass, the fluid-cell marking and the key / histogram of the P2G
// lists on the position it has just computed (the same device functions as k_collide_mark: identical results), so the
// positions cross HBM once instead of three times and the velocity is only written where a wall or the obstacle changed it.
// Row-walk unroll and register cap of the sweep, measured at S2 (both sweeps, ms) on the benchmark's near-lattice dam-break state:
// unroll 1 / 2 / 4 / 8 / 12 / 16 / 32 = 3.14 / 3.11 / 3.22 / 2.96 / 3.05 / 2.90* / 2.98; with the kernel held to 32 registers (8 CTAs of
// 256 threads per SM, a 16-byte spill) unroll 2 / 4 / 8 = 3.08 / 3.19 / 2.88 (*: together with the cap).  Without any minimum-blocks
// hint ptxas picks 32 / 39 registers, with a hint of 1 it takes 48 / 51 and the sweeps cost 3.76 ms.  The choice is state dependent:
// on the developed-flow state of bench.py (crowded cells, ragged rows) unroll 8 costs 4.32 ms against 4.02 (unroll 4) / 4.04 (unroll 2)
// under the same cap -- the longer unrolled groups run more masked-off candidates; the defaults follow the headline configuration.
#ifndef SEP_UNROLL
#define SEP_UNROLL 8
#endif
#ifndef SEP_MINB
#define SEP_MINB 8
#endif
constexpr int kSepUnroll = SEP_UNROLL, kSepMinBlocks = SEP_MINB;
template <bool COLORS, bool FUSE>
__global__ void __launch_bounds__(kThreads, kSepMinBlocks > 0 ? kSepMinBlocks : 1) k_separate(Geo g, const int* __restrict__ first, const float2* __restrict__ pos_in,
                                                       float2* __restrict__ pos_out, const float* __restrict__ col_in,
                                                       float* __restrict__ col_out, float2* __restrict__ vel, float ox, float oy,
                                                       float ovx, float ovy, float orad, int* __restrict__ cell_type,
                                                       int* __restrict__ gkeys, int* __restrict__ gcount) {
	const CollideArgs ca = collide_args(g, ox, oy, ovx, ovy, orad);
	const float min_dist = 2.f * g.r;
	const float min_dist_sq = min_dist * min_dist;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		float2 p = pos_in[t];
		float ax = p.x, ay = p.y;
		float c0 = 0.f, c1 = 0.f, c2 = 0.f;
		if (COLORS) { c0 = col_in[3 * (size_t)t]; c1 = col_in[3 * (size_t)t + 1]; c2 = col_in[3 * (size_t)t + 2]; }
		int pxi = (int)(p.x * g.p_inv), pyi = (int)(p.y * g.p_inv);
		int x0 = max(pxi - 1, 0), y0 = max(pyi - 1, g.hy0);
		int x1 = min(pxi + 1, g.pnx - 1), y1 = min(pyi + 1, g.hy1 - 1);
		// the (up to) three window rows are three contiguous slot ranges; candidates are numbered through them
		int ra[3], rn[3];
#pragma unroll
		for (int r = 0; r < 3; r++) {
			int yi = y0 + r;
			bool on = x0 <= x1 && yi <= y1;
			ra[r] = on ? first[x0 + g.pnx * yi] : 0;
			rn[r] = on ? first[x1 + g.pnx * yi + 1] - ra[r] : 0;
		}
		// Two passes keep the warp converged: pass 1 is the cheap distance test and leaves a bit mask of overlapping
		// partners; pass 2 runs the sqrt/divide push only for the set bits, in candidate order (the order of the adds is
		// part of the schedule: oracle orcflip_separate_jacobi).  A particle is its own candidate and drops out through
		// d2 == 0 like any coincident pair (flip_fluid.h:226).
		auto push = [&](int j) {
			float2 q = pos_in[j];
			float dx = q.x - p.x, dy = q.y - p.y;
			float d2 = dx * dx + dy * dy;
			float d = sqrtf(d2);
			float sc = .5f * (min_dist - d) / d;
			ax -= dx * sc;
			ay -= dy * sc;
			if (COLORS) {  // colour diffusion toward the pair mean, flip_fluid.h:239-245
				float q0 = col_in[3 * (size_t)j], q1 = col_in[3 * (size_t)j + 1], q2 = col_in[3 * (size_t)j + 2];
				c0 += .001f * ((c0 + q0) / 2.f - c0);
				c1 += .001f * ((c1 + q1) / 2.f - c1);
				c2 += .001f * ((c2 + q2) / 2.f - c2);
			}
		};
		if (max(rn[0], max(rn[1], rn[2])) <= 32) {
			// the usual window: each row is walked with a plain pointer (four loads in flight), its hits pushed before the next row
#pragma unroll
			for (int r = 0; r < 3; r++) {
				const float2* __restrict__ row = pos_in + ra[r];
				unsigned hits = 0;
#pragma unroll kSepUnroll
				for (int i = 0; i < rn[r]; i++) {
					float2 q = row[i];
					float dx = q.x - p.x, dy = q.y - p.y;
					float d2 = dx * dx + dy * dy;
					hits |= (unsigned)!(d2 > min_dist_sq || d2 == 0.f) << i;
				}
				while (hits) {
					int i = __ffs(hits) - 1;
					hits &= hits - 1;
					push(ra[r] + i);
				}
			}
		} else {
			// crowded window: candidates numbered through the three rows, chunks of 32; candidate k is slot k + b, b picked by
			// two compares (kept as selects: no branch in the candidate loop)
			const int n_all = rn[0] + rn[1] + rn[2];
			const int n0 = rn[0], n01 = rn[0] + rn[1];
			const int b0 = ra[0], b1 = ra[1] - n0, b2 = ra[2] - n01;
#pragma unroll 1
			for (int base = 0; base < n_all; base += 32) {
				const int n = min(32, n_all - base);
				unsigned hits = 0;
				for (int i = 0; i < n; i++) {
					int k = base + i;
					float2 q = pos_in[k + (k < n0 ? b0 : (k < n01 ? b1 : b2))];
					float dx = q.x - p.x, dy = q.y - p.y;
					float d2 = dx * dx + dy * dy;
					if (!(d2 > min_dist_sq || d2 == 0.f)) hits |= 1u << i;
				}
				while (hits) {
					int k = base + __ffs(hits) - 1;
					hits &= hits - 1;
					push(k + (k < n0 ? b0 : (k < n01 ? b1 : b2)));
				}
			}
		}
		float2 pn = make_float2(ax, ay);
		if (FUSE) {
			float2 v = vel[t];
			if (collide_one(ca, pn, v)) vel[t] = v;
			mark_one(g, pn, t, cell_type, gkeys, gcount);
		}
		pos_out[t] = pn;
		if (COLORS) { col_out[3 * (size_t)t] = c0; col_out[3 * (size_t)t + 1] = c1; col_out[3 * (size_t)t + 2] = c2; }
	}
}

// ------------------------------------------------------------------ F5 prep: cell_type from s (flip_fluid.h:348-350)
__global__ void __launch_bounds__(kThreads) k_cell_type_from_s(int n, const float* __restrict__ s, int* __restrict__ cell_type) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
		cell_type[c] = s[c] == 0.f ? SOLID_CELL : AIR_CELL;
}

// Is every s value 0 or 1?  The blocked solvers pack s into one bit per neighbour, which is exact only then; any other
// value raises the flag and the solve takes the general kernels.  Runs behind every write of s that can introduce such
// values (flip_upload / flip_upload_rows), on the device: no host scan, no host sync on the upload path.
__global__ void __launch_bounds__(kThreads) k_check_s_binary(size_t n, const float* __restrict__ s, int* __restrict__ flag) {
	bool bad = false;
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
		float v = s[c];
		bad |= !(v == 0.f || v == 1.f);
	}
	if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *flag = 1;  // racing writers store the same value
}

// One word device -> pinned host memory, written by a kernel: a cudaMemcpyAsync D2H of 4 bytes on the simulation stream
// queues on the D2H copy engine BEHIND a frame that is crossing PCIe (9 ms at S2) and stalls every kernel behind it.
__global__ void k_word_to_host(const int* __restrict__ d_word, volatile int* __restrict__ h_word) {
	if (threadIdx.x == 0 && blockIdx.x == 0) *h_word = *d_word;
}

// ------------------------------------------------------------------ F4 + F5 marking
// handleParticleCollisions (flip_fluid.h:254-289) fused with the fluid-cell marking of :352-359 and with
// the key + histogram of the second (index-only) counting sort: particles by the BASE CELL (x0,y0) of their
// bilinear stencil, taken from the final positions of this step.  That list is what makes the P2G a gather.
template <bool COLLIDE>
__global__ void __launch_bounds__(kThreads) k_collide_mark(Geo g, float2* __restrict__ pos, float2* __restrict__ vel, float ox,
                                                           float oy, float ovx, float ovy, float orad,
                                                           int* __restrict__ cell_type, int* __restrict__ gkeys,
                                                           int* __restrict__ gcount) {
	const CollideArgs ca = collide_args(g, ox, oy, ovx, ovy, orad);
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		float2 p = pos[t];
		if (COLLIDE) {
			float2 v = vel[t];
			collide_one(ca, p, v);
			pos[t] = p;
			vel[t] = v;
		}
		if (cell_type) mark_one(g, p, t, cell_type, gkeys, gcount);
	}
}

// fill of the index-only sort (slots claimed in any order) ...
__global__ void __launch_bounds__(kThreads) k_fill_grid(int np, const int* __restrict__ gkeys, int* __restrict__ gfirst,
                                                        int* __restrict__ gidx) {
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < np; t += gridDim.x * blockDim.x)
		gidx[atomicSub(gfirst + gkeys[t], 1) - 1] = t;
}
// ... then every cell's list is sorted by slot number, which fixes the order of the float adds of the gather
__global__ void __launch_bounds__(kThreads) k_fixup_grid(int n_cells, const int* __restrict__ gcount,
                                                         const int* __restrict__ gfirst, int* __restrict__ gidx, CrowdQueue crowd) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
		int n = gcount[c];
		if (n < 2) continue;
		if (n > kCrowd && crowd_push(crowd, c)) continue;
		int b = gfirst[c];
		if (n <= 16) {
			int v[16];
#pragma unroll
			for (int i = 0; i < 16; i++) v[i] = i < n ? gidx[b + i] : 0x7fffffff;
			sort_network_ascending<16>(v);
#pragma unroll
			for (int i = 0; i < 16; i++)
				if (i < n) gidx[b + i] = v[i];
			continue;
		}
		for (int i = 1; i < n; i++) {
			int v = gidx[b + i], j = i - 1;
			while (j >= 0 && gidx[b + j] > v) { gidx[b + j + 1] = gidx[b + j]; j--; }
			gidx[b + j + 1] = v;
		}
	}
}

// ------------------------------------------------------------------ F6 + F7: sort-ordered, atomic-free P2G
// transferVelocities(true) splat + normalise + solid restore (flip_fluid.h:362-403,432-446) and
// updateParticleDensity (:292-319) in GATHER form: one thread per grid node reads the particle lists of the
// (up to) four base cells whose stencils reach it -- cells (i-1..i, j-1..j); the two cells of a row are one
// contiguous range of the list -- and accumulates its own node in list order: no atomics, run-to-run
// bit-stable.  The per-particle arithmetic is the reference's; only the order of the adds into a node
// differs (slot order instead of caller order).  du, dv and particle_density are one array here: the
// reference computes the same weights three times (bit-equal except aliased border cells).
template <bool FAST>
__global__ void __launch_bounds__(256) k_p2g_gather(Geo g, const int* __restrict__ gfirst, const int* __restrict__ gidx,
                                                    const float2* __restrict__ pos, const float2* __restrict__ vel,
                                                    const int* __restrict__ cell_type, float* __restrict__ u,
                                                    float* __restrict__ v, float* __restrict__ prev_u,
                                                    float* __restrict__ prev_v, float* __restrict__ dens,
                                                    float* __restrict__ p) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int j = g.ny0 + blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx || j >= g.ny1) return;
	float su = 0.f, sv = 0.f, sw = 0.f;
	int cx0 = max(i - 1, 0);
	if (FAST && i >= 1 && i <= g.nx - 3 && j - 1 > g.gy0 && j < g.gy1 - 1 && j <= g.ny - 3) {
		// Interior node: the four base cells (i-1..i, j-1..j) are all left of / below the aliased last column / row
		// (Q13), so every particle listed under cell (cx, cy) has x0 == cx, x1 == cx+1, y0 == cy, y1 == cy+1 (the list
		// key IS (x0, y0), k_collide_mark; the edge rows of a shard's window, which also collect strays, are excluded)
		// and reaches this node with exactly one weight, known from the cell alone:
		// no index conversion, no match tests.  Same expressions and the same order of adds as the general loop.
#pragma unroll
		for (int r = 0; r < 2; r++) {
			const int cy = j - 1 + r;
			const int a = gfirst[(i - 1) + g.nx * cy], mid = gfirst[i + g.nx * cy], b = gfirst[i + g.nx * cy + 1];
			const float fy = (float)cy;
			for (int k = a; k < b; k++) {
				int t = gidx[k];
				float2 pp = pos[t];
				float2 pv = vel[t];
				const bool own = k >= mid;  // listed under cell i (weight sx) or under cell i-1 (weight tx)
				float x = clampf(pp.x, g.h, g.xmax_c), y = clampf(pp.y, g.h, g.ymax_c);
				float xs = x - g.h2, ys = y - g.h2;
				float tx = g.inv_h * (xs - g.h * (float)(own ? i : i - 1));
				float ty = g.inv_h * (ys - g.h * fy);
				float wx = own ? 1.f - tx : tx;
				float wy = r == 1 ? 1.f - ty : ty;
				float w = wx * wy;
				su += pv.x * w; sv += pv.y * w; sw += w;
			}
		}
	} else
	for (int cy = max(j - 1, g.gy0); cy <= j; cy++) {
		int a = gfirst[cx0 + g.nx * cy], b = gfirst[i + g.nx * cy + 1];
		for (int k = a; k < b; k++) {
			int t = gidx[k];
			float2 pp = pos[t];
			Stencil st = make_stencil(g, pp.x, pp.y);
			bool mx0 = st.x0 == i, mx1 = st.x1 == i, my0 = st.y0 == j, my1 = st.y1 == j;
			if (!((mx0 | mx1) & (my0 | my1))) continue;
			float2 pv = vel[t];
			if (mx0 & my0) { su += pv.x * st.w0; sv += pv.y * st.w0; sw += st.w0; }
			if (mx1 & my0) { su += pv.x * st.w1; sv += pv.y * st.w1; sw += st.w1; }
			if (mx1 & my1) { su += pv.x * st.w2; sv += pv.y * st.w2; sw += st.w2; }
			if (mx0 & my1) { su += pv.x * st.w3; sv += pv.y * st.w3; sw += st.w3; }
		}
	}
	int c = i + g.nx * j;
	if (sw > 0.f) { su /= sw; sv /= sw; }  // :433-435
	// restore solid cells, :437-445 (prev_u/prev_v of :340-341 == the values still in u/v here)
	int ct = cell_type[c];
	bool solid = ct == SOLID_CELL;
	if (solid || (i > 0 && cell_type[c - 1] == SOLID_CELL)) su = u[c];
	if (solid || (j > 0 && cell_type[c - g.nx] == SOLID_CELL)) sv = v[c];
	u[c] = su; v[c] = sv;
	prev_u[c] = su; prev_v[c] = sv;  // the copy solveIncompressibility makes at :453-454
	dens[c] = sw;
	p[c] = 0.f;                     // :452
}

// ------------------------------------------------------------------ F7 latch (flip_fluid.h:321-332, Q14)
__global__ void __launch_bounds__(256) k_rest_partials(int n, const int* __restrict__ cell_type, const float* __restrict__ dens,
                                                       const float* __restrict__ rest, float* __restrict__ partials) {
	if (*rest != 0.f) return;
	float s = 0.f, cnt = 0.f;
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
		if (cell_type[c] == FLUID_CELL) { s += dens[c]; cnt += 1.f; }
	for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
	__shared__ float ws[2][8];
	if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = s; ws[1][threadIdx.x >> 5] = cnt; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int i = 1; i < 8; i++) { s += ws[0][i]; cnt += ws[1][i]; }
		partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = cnt;
	}
}
__global__ void k_rest_latch(int n_partials, const float* __restrict__ partials, float* __restrict__ rest) {
	if (threadIdx.x != 0 || *rest != 0.f) return;
	double s = 0.0, cnt = 0.0;
	for (int i = 0; i < n_partials; i++) { s += partials[2 * i]; cnt += partials[2 * i + 1]; }
	if (cnt > 0.0) *rest = (float)(s / cnt);
}

// ------------------------------------------------------------------ F8: G2P (flip_fluid.h:404-429)
// PIC/FLIP blend gather.  Every gather is issued before the first use (one memory round trip after the position
// instead of three dependent ones: 1.28 -> 0.77 ms at S2), and the -x / -y neighbour types of the validity test
// :406-409 are taken from the stencil's own nodes where they coincide:
// cell_type[n1-1] is node 0 and cell_type[n2-1] is node 3 whenever x1 == x0+1, cell_type[n3-nx] is node 0 and
// cell_type[n2-nx] is node 1 whenever y1 == y0+1 (they differ only in the aliased last column / row, Q13).
// STAGE: the frame pipeline's caller-order copy of the (final) positions is written from here -- the kernel is bound by
// gather latency, not bandwidth, so the scattered 8-byte stores ride along (a separate un-permute pass costs 1.46 ms at S2).
// (five CTAs per SM: 48 registers and an 8-byte spill instead of 60 registers -- 0.76 -> 0.70 ms at S2; six CTAs: 0.74)
template <bool STAGE>
__global__ void __launch_bounds__(kThreads, 5) k_g2p(Geo g, const float2* __restrict__ pos, float2* __restrict__ vel,
                                                  const float* __restrict__ u, const float* __restrict__ v,
                                                  const float* __restrict__ prev_u, const float* __restrict__ prev_v,
                                                  const int* __restrict__ cell_type, float flip_ratio,
                                                  const int* __restrict__ orig, float2* __restrict__ staging,
                                                  int* __restrict__ vmax_bits) {
	float vmax = 0.f;  // sharded steps: largest |v_y| leaving this step, which sizes the next step's ghost band (k_shard_band)
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		float2 p = pos[t];
		if (STAGE) staging[orig[t]] = p;
		float2 pv = vel[t];
		Stencil st = make_stencil(g, p.x, p.y);
		const int n0 = st.x0 + g.nx * st.y0, n1 = st.x1 + g.nx * st.y0, n2 = st.x1 + g.nx * st.y1, n3 = st.x0 + g.nx * st.y1;
		const int ct0 = cell_type[n0], ct1 = cell_type[n1], ct2 = cell_type[n2], ct3 = cell_type[n3];
		const float fu0 = u[n0], fu1 = u[n1], fu2 = u[n2], fu3 = u[n3];
		const float fv0 = v[n0], fv1 = v[n1], fv2 = v[n2], fv3 = v[n3];
		const float pu0 = prev_u[n0], pu1 = prev_u[n1], pu2 = prev_u[n2], pu3 = prev_u[n3];
		const float pw0 = prev_v[n0], pw1 = prev_v[n1], pw2 = prev_v[n2], pw3 = prev_v[n3];
		// Q3: the reference reads cell_type[n-off] below index 0 when y0==0; policy: out of range => not fluid (-1)
		const bool xadj = st.x1 == st.x0 + 1, yadj = st.y1 == st.y0 + 1;
		const int l0 = n0 - 1 >= 0 ? cell_type[n0 - 1] : -1;
		const int l3 = n3 - 1 >= 0 ? cell_type[n3 - 1] : -1;
		const int l1 = xadj ? ct0 : (n1 - 1 >= 0 ? cell_type[n1 - 1] : -1);
		const int l2 = xadj ? ct3 : (n2 - 1 >= 0 ? cell_type[n2 - 1] : -1);
		const int b0 = n0 - g.nx >= 0 ? cell_type[n0 - g.nx] : -1;
		const int b1 = n1 - g.nx >= 0 ? cell_type[n1 - g.nx] : -1;
		const int b3 = yadj ? ct0 : (n3 - g.nx >= 0 ? cell_type[n3 - g.nx] : -1);
		const int b2 = yadj ? ct1 : (n2 - g.nx >= 0 ? cell_type[n2 - g.nx] : -1);
#pragma unroll
		for (int comp = 0; comp < 2; comp++) {
			const int m0 = comp == 0 ? l0 : b0, m1 = comp == 0 ? l1 : b1, m2 = comp == 0 ? l2 : b2, m3 = comp == 0 ? l3 : b3;
			const float f0 = comp == 0 ? fu0 : fv0, f1 = comp == 0 ? fu1 : fv1, f2 = comp == 0 ? fu2 : fv2, f3 = comp == 0 ? fu3 : fv3;
			const float q0 = comp == 0 ? pu0 : pw0, q1 = comp == 0 ? pu1 : pw1, q2 = comp == 0 ? pu2 : pw2, q3 = comp == 0 ? pu3 : pw3;
			float k0 = (ct0 == FLUID_CELL || m0 == FLUID_CELL) ? 1.f : 0.f;
			float k1 = (ct1 == FLUID_CELL || m1 == FLUID_CELL) ? 1.f : 0.f;
			float k2 = (ct2 == FLUID_CELL || m2 == FLUID_CELL) ? 1.f : 0.f;
			float k3 = (ct3 == FLUID_CELL || m3 == FLUID_CELL) ? 1.f : 0.f;
			float dsum = k0 * st.w0 + k1 * st.w1 + k2 * st.w2 + k3 * st.w3;
			if (dsum > 0.f) {
				float pic = (k0 * st.w0 * f0 + k1 * st.w1 * f1 + k2 * st.w2 * f2 + k3 * st.w3 * f3) / dsum;
				float corr = (k0 * st.w0 * (f0 - q0) + k1 * st.w1 * (f1 - q1) + k2 * st.w2 * (f2 - q2) +
				              k3 * st.w3 * (f3 - q3)) / dsum;
				float cur = comp == 0 ? pv.x : pv.y;
				float flipv = cur + corr;
				float nv = (1.f - flip_ratio) * pic + flip_ratio * flipv;
				if (comp == 0) pv.x = nv; else pv.y = nv;
			}
		}
		vel[t] = pv;
		vmax = fmaxf(vmax, fabsf(pv.y));
	}
	if (vmax_bits) {
		const int mi = warp_max(__float_as_int(vmax));
		if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(vmax_bits, mi);
	}
}

// ------------------------------------------------------------------ F10: render state (flip_fluid.h:498-557)
__global__ void __launch_bounds__(kThreads) k_particle_colors(Geo g, const float2* __restrict__ pos, float* __restrict__ col,
                                                              const float* __restrict__ dens, const float* __restrict__ rest_p) {
	float rest = *rest_p;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		float r = clampf(col[3 * (size_t)t] - .01f, 0.f, 1.f);
		float gg = clampf(col[3 * (size_t)t + 1] - .01f, 0.f, 1.f);
		float b = clampf(col[3 * (size_t)t + 2] + .01f, 0.f, 1.f);
		if (rest > 0.f) {
			float2 p = pos[t];
			int xi = clampi((int)(g.inv_h * p.x), 1, g.nx - 1);
			int yi = clampi((int)(g.inv_h * p.y), 1, g.ny - 1);
			float rel = dens[xi + g.nx * yi] / rest;
			if (rel < .7f) { r = .8f; gg = .8f; b = 1.f; }
		}
		col[3 * (size_t)t] = r; col[3 * (size_t)t + 1] = gg; col[3 * (size_t)t + 2] = b;
	}
}
__global__ void __launch_bounds__(kThreads) k_cell_colors(int n, const int* __restrict__ cell_type, const float* __restrict__ dens,
                                                          const float* __restrict__ rest_p, float* __restrict__ cc) {
	float rest = *rest_p;
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
		float r = 0.f, gg = 0.f, b = 0.f;
		int ct = cell_type[c];
		if (ct == SOLID_CELL) { r = gg = b = .5f; }
		else if (ct == FLUID_CELL) {  // setSciColor(i, d, 0, 2), :521-540
			float val = dens[c];
			if (rest > 0.f) val /= rest;
			val = clampf(val, 0.f, 2.f - .0001f);
			float span = 2.f - 0.f;
			val = span == 0.f ? .5f : (val - 0.f) / span;
			float m = .25f;
			int num = (int)(val / m);
			float sc = (val - (float)num * m) / m;
			switch (num) {
				case 0: r = 0.f; gg = sc; b = 1.f; break;
				case 1: r = 0.f; gg = 1.f; b = 1.f - sc; break;
				case 2: r = sc; gg = 1.f; b = 0.f; break;
				case 3: r = 1.f; gg = 1.f - sc; b = 0.f; break;
			}
		}
		cc[3 * (size_t)c] = r; cc[3 * (size_t)c + 1] = gg; cc[3 * (size_t)c + 2] = b;
	}
}

// ------------------------------------------------------------------ setObstacle (flip_fluid/src/main.cpp:104-122, Q10)
__global__ void __launch_bounds__(256) k_set_obstacle(Geo g, float x, float y, float vx, float vy, float rad,
                                                      float* __restrict__ s, float* __restrict__ u, float* __restrict__ v) {
	int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
	int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx - 2 || j >= g.ny - 2) return;
	int c = i + g.nx * j;
	float dx = g.h * (.5f + (float)i) - x, dy = g.h * (.5f + (float)j) - y;
	float sv = 1.f;
	if (dx * dx + dy * dy < rad * rad) {
		sv = 0.f;
		u[c] = vx; u[c + 1] = vx; v[c] = vy; v[c + g.nx] = vy;  // duplicates across cells store the same value
	}
	s[c] = sv;
}


// ------------------------------------------------------------------ slab sharding: particle ownership and ghosts
// Ownership is by position: a particle belongs to the rank whose slab holds the grid row of its cell.  At the start
// of a step every rank keeps the particles it owns (out of everything it simulated last step, ghosts included: a
// particle that crossed the boundary was a valid ghost on its new owner) and copies the ones within `band` rows of a
// slab boundary into the neighbour's send buffer.  Slot order does not matter: the counting sort orders by
// (hash cell, global id), which is the same relative order on any partition.
struct ShardPack {
	float2* pos; float2* vel; int* id;
};
// largest |v_y| over the local particles, as raw float bits (non-negative floats order like ints)
__global__ void __launch_bounds__(kThreads) k_shard_vmax(int np, const float2* __restrict__ vel, int* __restrict__ vmax_bits) {
	float m = 0.f;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < np; t += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(vel[t].y));
	int mi = warp_max(__float_as_int(m));
	if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(vmax_bits, mi);
}
// Ghost band in rows: every particle that can touch this rank's rows during the step must be present.  A particle moves at
// most (vmax + |g| dt) dt by integration, then up to one overlap (2.2r hash cell + r push) per separation sweep -- each
// Jacobi sweep also widens the dependency cone by one hash cell -- and its bilinear stencil reaches one more cell; +2 rows
// of slack.  The same value on every rank (vmax is all-reduced); the host refuses a band wider than the thinnest slab.
__global__ void k_shard_band(int* __restrict__ vmax_bits, float dt, float gravity, float h, float r, int sep_iters, int* __restrict__ band) {
	if (threadIdx.x != 0) return;
	float v = __int_as_float(*vmax_bits) + fabsf(gravity) * dt;
	*vmax_bits = 0;  // consumed: the G2P of this step accumulates the next one
	float reach = v * dt / h + (float)sep_iters * (3.2f * r) / h;
	int b = (int)ceilf(reach) + 1 + 2;
	*band = max(b, 6);
}

// Step-start bookkeeping.  Ownership is by position: a rank owns the particles whose cell row lies in its slab.  Out of
// every copy it simulated last step (its own particles and the ghosts) a rank keeps exactly those; a particle that
// crossed the boundary was a ghost on its new owner, whose copy -- integrated, separated and G2P'd inside that rank's
// valid rows -- is the one that survives.  (The band below is wide enough that every such particle WAS a ghost there.)
// Kept particles within `band` rows of a boundary are copied to that neighbour as this step's ghosts.
// Ownership is by position: a particle belongs to the rank whose slab holds its cell row; owned particles within `band` rows of
// a slab boundary also go to that neighbour (as ghosts, or as migrants it now owns).
struct ShardClass { bool own, go_down, go_up; };
__device__ __forceinline__ ShardClass shard_class(const Geo& g, float2 p, int j0, int j1, int band, int has_down, int has_up) {
	const int row = clampi((int)(g.inv_h * p.y), 0, g.ny - 1);
	ShardClass c;
	c.own = row >= j0 && row < j1;
	c.go_down = c.own && has_down && row < j0 + band;
	c.go_up = c.own && has_up && row >= j1 - band;
	return c;
}
// Stable compaction of the owned particles (flip_readback_particles; boundary copies with `exchange`) in two passes over
// 256-particle chunks -- count, one scan of the chunk counts, place -- one WARP per chunk (no block barriers, the chunk's loads
// in flight together).  The step itself does not compact: see k_shard_key.
constexpr int kShardChunk = 256;
__global__ void __launch_bounds__(kThreads) k_shard_count(Geo g, int j0, int j1, const int* __restrict__ band_p, int has_down, int has_up,
                                                          const float2* __restrict__ pos, int n_chunks, int* __restrict__ chunk_cnt /* [3][n_chunks] */) {
	const int band = band_p ? *band_p : 0;
	const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
	for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps) {
		int n0 = 0, n1 = 0, n2 = 0;
#pragma unroll
		for (int i = 0; i < kShardChunk / 32; i++) {
			const int t = chunk * kShardChunk + i * 32 + lane;
			ShardClass c{false, false, false};
			if (t < g.np) c = shard_class(g, pos[t], j0, j1, band, has_down, has_up);
			n0 += __popc(__ballot_sync(0xffffffffu, c.own));
			n1 += __popc(__ballot_sync(0xffffffffu, c.go_down));
			n2 += __popc(__ballot_sync(0xffffffffu, c.go_up));
		}
		if (lane == 0) { chunk_cnt[chunk] = n0; chunk_cnt[n_chunks + chunk] = n1; chunk_cnt[2 * n_chunks + chunk] = n2; }
	}
}
// chunk_incl: inclusive running sum over the three count rows laid end to end (running_sum)
__global__ void __launch_bounds__(kThreads) k_shard_place(Geo g, int j0, int j1, const int* __restrict__ band_p, int has_down, int has_up,
                                                          const float2* __restrict__ pos, const float2* __restrict__ vel, const int* __restrict__ id,
                                                          ShardPack keep, ShardPack down, ShardPack up, int n_chunks,
                                                          const int* __restrict__ chunk_cnt, const int* __restrict__ chunk_incl, int cap_send,
                                                          int* __restrict__ counters) {
	const int band = band_p ? *band_p : 0;
	const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
	const unsigned lt = (1u << lane) - 1u;
	const int tot0 = chunk_incl[n_chunks - 1], tot1 = chunk_incl[2 * n_chunks - 1], tot2 = chunk_incl[3 * n_chunks - 1];
	if (blockIdx.x == 0 && threadIdx.x == 0) { counters[0] = tot0; counters[1] = tot1 - tot0; counters[2] = tot2 - tot1; }
	for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps) {
		int at0 = chunk_incl[chunk] - chunk_cnt[chunk];
		int at1 = chunk_incl[n_chunks + chunk] - chunk_cnt[n_chunks + chunk] - tot0;
		int at2 = chunk_incl[2 * n_chunks + chunk] - chunk_cnt[2 * n_chunks + chunk] - tot1;
#pragma unroll 4
		for (int i = 0; i < kShardChunk / 32; i++) {
			const int t = chunk * kShardChunk + i * 32 + lane;
			float2 p = make_float2(0.f, 0.f), v = make_float2(0.f, 0.f);
			int gid = 0;
			ShardClass c{false, false, false};
			if (t < g.np) {
				p = pos[t];
				c = shard_class(g, p, j0, j1, band, has_down, has_up);
				if (c.own) { v = vel[t]; gid = id[t]; }
			}
			const unsigned m0 = __ballot_sync(0xffffffffu, c.own), m1 = __ballot_sync(0xffffffffu, c.go_down), m2 = __ballot_sync(0xffffffffu, c.go_up);
			// boundary copies go out as they are: their receiver integrates them
			if (c.go_down) { const int k = at1 + __popc(m1 & lt); if (k < cap_send) { down.pos[k] = p; down.vel[k] = v; down.id[k] = gid; } }
			if (c.go_up) { const int k = at2 + __popc(m2 & lt); if (k < cap_send) { up.pos[k] = p; up.vel[k] = v; up.id[k] = gid; } }
			if (c.own) { const int k = at0 + __popc(m0 & lt); keep.pos[k] = p; keep.vel[k] = v; keep.id[k] = gid; }
			at0 += __popc(m0); at1 += __popc(m1); at2 += __popc(m2);
		}
	}
}
// Step start of a sharded step, one pass over what the rank holds (owned particles and last step's ghosts, in hash-cell order):
// owned particles get this step's sort key from their integrated position (the reorder pass of the sort redoes the integration,
// as in the one-GPU step: k_integrate_bin STORE = false) and enter the hash histogram; everything else gets kDeadKey and drops
// out of the sort, so no compaction pass is needed; owned particles within `band` rows of a slab boundary are also copied, as they
// are, into the send buffer of that neighbour (slots claimed per warp: only the warps at a boundary have any).
constexpr int kDeadKey = -1;
__global__ void __launch_bounds__(kThreads) k_shard_key(Geo g, int j0, int j1, const int* __restrict__ band_p, int has_down, int has_up,
                                                        const float2* __restrict__ pos, const float2* __restrict__ vel, const int* __restrict__ id,
                                                        float dt, float gravity, int* __restrict__ keys, int* __restrict__ count,
                                                        ShardPack down, ShardPack up, int cap_send, int* __restrict__ counters) {
	const int band = *band_p;
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int n_own = 0;
	for (int t0 = blockIdx.x * blockDim.x; t0 < g.np; t0 += gridDim.x * blockDim.x) {   // (whole warps stay in the loop: ballots below)
		const int t = t0 + threadIdx.x;
		ShardClass c{false, false, false};
		float2 p = make_float2(0.f, 0.f), v = make_float2(0.f, 0.f);
		if (t < g.np) {
			p = pos[t];
			c = shard_class(g, p, j0, j1, band, has_down, has_up);
			if (c.own) {
				v = vel[t];
				float2 pi = p, vi = v;
				integrate_one(pi, vi, dt, gravity);
				const int key = hash_cell(g, pi.x, pi.y);
				keys[t] = key;
				atomicAdd(count + key, 1);
				n_own++;
			} else {
				keys[t] = kDeadKey;
			}
		}
		const unsigned m1 = __ballot_sync(0xffffffffu, c.go_down), m2 = __ballot_sync(0xffffffffu, c.go_up);
		if (m1) {
			int base = 0;
			if (lane == 0) base = atomicAdd(counters + 1, __popc(m1));
			base = __shfl_sync(0xffffffffu, base, 0);
			const int k = base + __popc(m1 & lt);
			if (c.go_down && k < cap_send) { down.pos[k] = p; down.vel[k] = v; down.id[k] = id[t]; }
		}
		if (m2) {
			int base = 0;
			if (lane == 0) base = atomicAdd(counters + 2, __popc(m2));
			base = __shfl_sync(0xffffffffu, base, 0);
			const int k = base + __popc(m2 & lt);
			if (c.go_up && k < cap_send) { up.pos[k] = p; up.vel[k] = v; up.id[k] = id[t]; }
		}
	}
	// how many this rank keeps: one atomic per CTA
	for (int o = 16; o > 0; o >>= 1) n_own += __shfl_xor_sync(0xffffffffu, n_own, o);
	__shared__ int s_own[kThreads / 32];
	if (lane == 0) s_own[threadIdx.x >> 5] = n_own;
	__syncthreads();
	if (threadIdx.x == 0) {
		int tot = 0;
		for (int w = 0; w < kThreads / 32; w++) tot += s_own[w];
		if (tot) atomicAdd(counters, tot);
	}
}
__global__ void k_shard_check(int* __restrict__ c, int cap_send, int held, int max_particles) {
	if (threadIdx.x != 0) return;
	// [1], [2]: boundary copies leaving; [3], [4]: arriving behind the `held` particles already in the buffer
	c[7] = (c[1] > cap_send || c[2] > cap_send || (long long)held + c[3] + c[4] > (long long)max_particles) ? 1 : 0;
}

// rest-density latch across slabs: per-rank (sum, count) over own rows, summed over ranks (flip_fluid.h:321-332)
__global__ void k_rest_pair(int n_partials, const float* __restrict__ partials, const float* __restrict__ rest, float* __restrict__ pair) {
	if (threadIdx.x != 0) return;
	double s = 0.0, cnt = 0.0;
	if (*rest == 0.f)
		for (int i = 0; i < n_partials; i++) { s += partials[2 * i]; cnt += partials[2 * i + 1]; }
	pair[0] = (float)s; pair[1] = (float)cnt;
}
__global__ void k_rest_from_pair(const float* __restrict__ pair, float* __restrict__ rest) {
	if (threadIdx.x == 0 && *rest == 0.f && pair[1] > 0.f) *rest = pair[0] / pair[1];
}


// =================================================================== reference-order ("exact") mode
// flip_set_exact_order(sim, 1): every place where the production path only changes the ORDER of floating-point
// operations is switched to the reference's own order, so that whole steps are bit-identical to the serial CPU code
// (with the lexicographic solver).  Meant for parity runs on small scenes; it is much slower.

// ---- F3 in the reference's serial order (flip_fluid.h:203-250), made parallel by level scheduling.
// Processing particle i reads/writes only particles of the 3x3 hash cells around its current cell, so two particles
// whose bins are more than kDepR cells apart (Chebyshev) commute; level[i] = 1 + max level of the lower-indexed particles
// within kDepR cells.  Running the levels in order, and the particles of one level concurrently, performs exactly the
// serial sweep.  (kDepR = 4 assumes no particle has drifted more than one cell from its bin, which k_sep_exact checks.)
constexpr int kDepR = 4;
__global__ void __launch_bounds__(kThreads) k_sep_levels(Geo g, const int* __restrict__ first, const int* __restrict__ orig,
                                                         const float2* __restrict__ binpos, const int* __restrict__ lvl_in,
                                                         int* __restrict__ lvl_out, int* __restrict__ changed) {
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		int cx, cy;
		hash_cell(g, binpos[t].x, binpos[t].y, &cx, &cy);
		const int me = orig[t];
		int lv = 0;
		for (int yi = max(cy - kDepR, g.hy0); yi <= min(cy + kDepR, g.hy1 - 1); yi++) {
			int a = first[max(cx - kDepR, 0) + g.pnx * yi], b = first[min(cx + kDepR, g.pnx - 1) + g.pnx * yi + 1];
			for (int j = a; j < b; j++)
				if (orig[j] < me) lv = max(lv, lvl_in[j]);
		}
		lv += 1;
		lvl_out[t] = lv;
		if (lv != lvl_in[t]) *changed = 1;
	}
}
template <bool COLORS>
__global__ void __launch_bounds__(kThreads) k_sep_exact(Geo g, const int* __restrict__ first, const int* __restrict__ level, int cur_level,
                                                        float2* __restrict__ pos, float* __restrict__ col, const float2* __restrict__ binpos,
                                                        int* __restrict__ drift_flag) {
	const float min_dist = 2.f * g.r;
	const float min_dist_sq = min_dist * min_dist;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.np; t += gridDim.x * blockDim.x) {
		if (level[t] != cur_level) continue;
		float px = pos[t].x, py = pos[t].y;
		int pxi = (int)(px * g.p_inv), pyi = (int)(py * g.p_inv);
		{	// the level schedule is only valid while particles stay within one cell of their bin
			int bx, by;
			hash_cell(g, binpos[t].x, binpos[t].y, &bx, &by);
			if (abs(clampi(pxi, 0, g.pnx - 1) - bx) > 1 || abs(clampi(pyi, g.hy0, g.hy1 - 1) - by) > 1) *drift_flag = 1;
		}
		int x0 = max(pxi - 1, 0), y0 = max(pyi - 1, g.hy0);
		int x1 = min(pxi + 1, g.pnx - 1), y1 = min(pyi + 1, g.hy1 - 1);
		for (int xi = x0; xi <= x1; xi++)
			for (int yi = y0; yi <= y1; yi++) {
				int c = xi + g.pnx * yi;
				for (int j = first[c]; j < first[c + 1]; j++) {
					if (j == t) continue;
					float dx = pos[j].x - px, dy = pos[j].y - py;
					float d2 = dx * dx + dy * dy;
					if (d2 > min_dist_sq || d2 == 0.f) continue;
					float d = sqrtf(d2);
					float sc = .5f * (min_dist - d) / d;
					dx *= sc; dy *= sc;
					pos[t].x -= dx;       // :234-235: i moves in memory ...
					pos[t].y -= dy;
					pos[j].x += dx;       // :236 (Q1): the partner only in x,
					py += dy;             //            the local py drifts instead
					if (COLORS)
						for (int ch = 0; ch < 3; ch++) {  // :239-245
							float c0 = col[ch + 3 * (size_t)t], c1 = col[ch + 3 * (size_t)j];
							float mean = (c0 + c1) / 2.f;
							col[ch + 3 * (size_t)t] = c0 + .001f * (mean - c0);
							col[ch + 3 * (size_t)j] = c1 + .001f * (mean - c1);
						}
				}
			}
	}
}

// ---- F6/F7 with the adds of every node in CALLER order: the lists of the four base cells are merged by caller index
__global__ void __launch_bounds__(kThreads) k_fixup_grid_by_id(int n_cells, const int* __restrict__ gcount, const int* __restrict__ gfirst,
                                                               int* __restrict__ gidx, const int* __restrict__ orig) {
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
		int n = gcount[c];
		if (n < 2) continue;
		int b = gfirst[c];
		for (int i = 1; i < n; i++) {
			int v = gidx[b + i], key = orig[v], j = i - 1;
			while (j >= 0 && orig[gidx[b + j]] > key) { gidx[b + j + 1] = gidx[b + j]; j--; }
			gidx[b + j + 1] = v;
		}
	}
}
__global__ void __launch_bounds__(256) k_p2g_gather_exact(Geo g, const int* __restrict__ gfirst, const int* __restrict__ gidx,
                                                          const int* __restrict__ orig, const float2* __restrict__ pos,
                                                          const float2* __restrict__ vel, const int* __restrict__ cell_type,
                                                          float* __restrict__ u, float* __restrict__ v, float* __restrict__ prev_u,
                                                          float* __restrict__ prev_v, float* __restrict__ dens, float* __restrict__ p) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int j = g.ny0 + blockIdx.y * blockDim.y + threadIdx.y;
	if (i >= g.nx || j >= g.ny1) return;
	// heads of the (up to) four base-cell lists that can reach this node
	int head[4], end[4];
	int nl = 0;
	for (int cy = max(j - 1, g.gy0); cy <= j; cy++)
		for (int cx = max(i - 1, 0); cx <= i; cx++) { head[nl] = gfirst[cx + g.nx * cy]; end[nl] = gfirst[cx + g.nx * cy + 1]; nl++; }
	float su = 0.f, sv = 0.f, swu = 0.f, swd = 0.f;
	for (;;) {
		int best = -1, best_id = 0x7fffffff;
		for (int l = 0; l < nl; l++)
			if (head[l] < end[l]) { int id = orig[gidx[head[l]]]; if (id < best_id) { best_id = id; best = l; } }
		if (best < 0) break;
		int t = gidx[head[best]++];
		float2 pp = pos[t];
		Stencil st = make_stencil(g, pp.x, pp.y);
		bool mx0 = st.x0 == i, mx1 = st.x1 == i, my0 = st.y0 == j, my1 = st.y1 == j;
		if (!((mx0 | mx1) & (my0 | my1))) continue;
		float2 pv = vel[t];
		// velocity splat order of :400-403 (nr0, nr1, nr2, nr3) ...
		if (mx0 & my0) { su += pv.x * st.w0; sv += pv.y * st.w0; swu += st.w0; }
		if (mx1 & my0) { su += pv.x * st.w1; sv += pv.y * st.w1; swu += st.w1; }
		if (mx1 & my1) { su += pv.x * st.w2; sv += pv.y * st.w2; swu += st.w2; }
		if (mx0 & my1) { su += pv.x * st.w3; sv += pv.y * st.w3; swu += st.w3; }
		// ... and the density order of :315-318 (x0y0, x1y0, x0y1, x1y1), which differs on aliased border cells
		if (mx0 & my0) swd += st.w0;
		if (mx1 & my0) swd += st.w1;
		if (mx0 & my1) swd += st.w3;
		if (mx1 & my1) swd += st.w2;
	}
	int c = i + g.nx * j;
	if (swu > 0.f) { su /= swu; sv /= swu; }
	int ct = cell_type[c];
	bool solid = ct == SOLID_CELL;
	if (solid || (i > 0 && cell_type[c - 1] == SOLID_CELL)) su = u[c];
	if (solid || (j > 0 && cell_type[c - g.nx] == SOLID_CELL)) sv = v[c];
	u[c] = su; v[c] = sv;
	prev_u[c] = su; prev_v[c] = sv;
	dens[c] = swd;
	p[c] = 0.f;
}
// the latch as the serial running sum of :321-332
__global__ void k_rest_latch_serial(int n, const int* __restrict__ cell_type, const float* __restrict__ dens, float* __restrict__ rest) {
	if (threadIdx.x != 0 || blockIdx.x != 0 || *rest != 0.f) return;
	float sum = 0.f;
	int cnt = 0;
	for (int c = 0; c < n; c++)
		if (cell_type[c] == FLUID_CELL) { sum += dens[c]; cnt++; }
	if (cnt > 0) *rest = sum / cnt;
}

}  // namespace fg

// =================================================================== host side
using namespace fg;

struct FlipSim {
	FlipDims d{};
	int device = 0;
	cudaStream_t st = nullptr;
	float *u = nullptr, *v = nullptr, *prev_u = nullptr, *prev_v = nullptr, *p = nullptr, *s = nullptr, *dens = nullptr;
	float *u_alt = nullptr, *v_alt = nullptr, *p_alt = nullptr;  // ping-pong partners for the tiled red-black solve (lazy)
	int* cell_type = nullptr;
	float* cell_color = nullptr;  // lazily allocated
	float2* pos[2] = {nullptr, nullptr};
	float2* vel[2] = {nullptr, nullptr};
	float* col[2] = {nullptr, nullptr};  // lazily allocated
	int* orig[2] = {nullptr, nullptr};   // caller index per slot; orig[cur] doubles as cell_particle_ids
	int cur = 0;
	int *keys = nullptr, *src = nullptr, *count = nullptr, *first = nullptr, *tile_sum = nullptr;
	int* crowd_list = nullptr;   // crowded hash cells / grid cells of this step's sorts (k_fixup_crowded); [kCrowdCap] + 1 count word

	int *gcount = nullptr, *gfirst = nullptr;  // particles per stencil base cell (P2G gather lists; keys/src are reused for them)
	float* d_rest = nullptr;   // device scalar particle_rest_density
	float* d_partials = nullptr;
	float* h_pinned = nullptr;  // [0] rest mirror
	// asynchronous readback of particle_pos (render thread consumes frame k while frame k+1 is computed)
	// Frame pipeline.  Once the caller has asked for a frame, the last G2P kernel of every flip_step also writes the positions
	// in caller order into a staging buffer, and flip_readback_pos_async only queues the PCIe copy, on a second stream.
	// Two staging buffers: frame k+1 is staged while frame k is still crossing PCIe.  (Un-permuting on a side stream right
	// after the collision pass, concurrently with the pressure solve, was measured and dropped: the persistent solver kernel
	// lost 3.5 ms to the contention -- profiles/r2_e2e.md.)
	cudaStream_t st_copy = nullptr;
	cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
	// flip_upload_async of the grid fields the caller rewrites every frame (s, u, v): the H2D copy runs on its own stream,
	// behind everything queued so far (the readers of the old values), and the step joins it only where those fields are
	// first read (collision / marking pass) -- it hides behind integrate + binning + separation
	cudaStream_t st_up = nullptr;
	cudaEvent_t ev_fence = nullptr, ev_uploaded = nullptr;
	bool upload_pending = false;
	float2* staging[2] = {nullptr, nullptr};
	bool copy_pending[2] = {false, false};
	bool frame_mode = false;     // set by the first flip_readback_pos_async
	bool staged_valid = false;   // staging[frame_idx] holds the caller-order positions of the current state
	int frame_idx = 0;
	SorWorkspace* ws = nullptr;
	// slab sharding (flip_create_sharded): rank owns grid rows [j0, j1); arrays keep the global size, only the window is processed
	bool sharded = false;
	int rank = 0, nranks = 1, j0 = 0, j1 = 0, halo = 9, band = 6, min_slab_rows = 0;
	int gy0 = 0, gy1 = 0, hy0 = 0, hy1 = 0;
	ncclComm_t comm = nullptr;
	float2 *send_pos[2] = {nullptr, nullptr}, *send_vel[2] = {nullptr, nullptr};
	int* send_id[2] = {nullptr, nullptr};
	int cap_send = 0;
	int* d_chunk = nullptr;      // [3][chunks of 256 particles] per-chunk counts of the kept / down / up classes, then (at chunk_stride) their running sum
	size_t chunk_stride = 0;
	int* d_counters = nullptr;   // [0] keep [1] down [2] up [3] from-down [4] from-up [6] ghost band (rows) [7] error flag [8] vmax bits
	bool vmax_valid = false;     // d_counters[8] holds the largest |v_y| of the current particle state (left by the G2P of a sharded step)
	int global_particles = 0;
	int* h_counters = nullptr;   // pinned mirror
	int n_owned = 0;
	long long exchanges = 0;
	bool s_binary = true;       // every s[] is 0 or 1 (checked on the device behind every upload of s)
	bool s_check_pending = false;
	int* d_sflag = nullptr;     // device: set by k_check_s_binary
	int* h_sflag = nullptr;     // pinned mirror, valid once ev_scheck has completed
	cudaEvent_t ev_scheck = nullptr;
	bool exact_order = false;   // flip_set_exact_order: reference order of every float reduction (parity runs)
	int *lvl[2] = {nullptr, nullptr}, *d_flags = nullptr;  // level schedule of the serial-order separation
	int sep_levels = 0;
	bool identity = true;       // slot order == caller order
	bool bins_valid = false;
	bool rest_seen = false;
	bool timings_on = false;
	float ph_ms[FLIP_PH_COUNT] = {0};
	static constexpr int kEvRing = 32;  // per-step event sets; a slot is only waited on when it is reused 32 steps later
	cudaEvent_t ev[kEvRing][FLIP_PH_COUNT + 1] = {{nullptr}};
	bool ev_pending[kEvRing] = {false};
	int ev_slot = 0;
	LaunchCounter lc;
	// Whole-step CUDA graph (single GPU, reference solver order, steady state): one captured simulate() per parity of the
	// particle ping-pong buffers, replayed while nothing it depends on changes.  Small scenes are launch-bound (S0: ~25
	// launches of a few microseconds each), which is what this removes; large scenes run the same graph without loss.
	struct StepGraph {
		cudaGraphExec_t exec = nullptr;
		FlipStepParams sp{};
		int np = -1;
		bool s_binary = true;
		long long launches = 0;
	} graphs[2];
	int eager_steps = 0;          // eager steps with unchanged parameters since the last change (allocations, plans, latches settle)
	FlipStepParams last_sp{};
	bool graphs_on = true;        // FLIPGPU_GRAPH=0 switches the replay off (A/B timing)
	long long graph_replays = 0;
};

namespace {
constexpr int kPartials = 1024;

Geo geo(const FlipSim* f) {
	Geo g;
	g.nx = f->d.f_num_x; g.ny = f->d.f_num_y; g.pnx = f->d.p_num_x; g.pny = f->d.p_num_y; g.np = f->d.num_particles;
	g.h = f->d.h; g.inv_h = f->d.f_inv_spacing; g.h2 = .5f * f->d.h; g.r = f->d.particle_radius; g.p_inv = f->d.p_inv_spacing;
	g.xmax_c = f->d.h * (float)(f->d.f_num_x - 1);
	g.ymax_c = f->d.h * (float)(f->d.f_num_y - 1);
	if (f->sharded) { g.gy0 = f->gy0; g.gy1 = f->gy1; g.hy0 = f->hy0; g.hy1 = f->hy1; g.ny0 = f->j0; g.ny1 = f->j1; }
	else { g.gy0 = 0; g.gy1 = g.ny; g.hy0 = 0; g.hy1 = g.pny; g.ny0 = 0; g.ny1 = g.ny; }
	return g;
}
unsigned pgrid(const FlipSim* f) { return wave_grid((size_t)std::max(f->d.num_particles, 1), kThreads); }

int ensure_colors(FlipSim* f) {
	if (f->col[0]) return FLIPGPU_OK;
	size_t n = 3 * (size_t)std::max(f->d.max_particles, 1);
	FG_CUDA(cudaMalloc(&f->col[0], n * 4));
	FG_CUDA(cudaMalloc(&f->col[1], n * 4));
	k_init_colors<<<wave_grid(f->d.max_particles, kThreads), kThreads, 0, f->st>>>(f->d.max_particles, f->col[0]);
	k_init_colors<<<wave_grid(f->d.max_particles, kThreads), kThreads, 0, f->st>>>(f->d.max_particles, f->col[1]);
	f->lc.n += 2;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}
int ensure_cell_colors(FlipSim* f) {
	if (f->cell_color) return FLIPGPU_OK;
	FG_CUDA(cudaMalloc(&f->cell_color, 3 * (size_t)f->d.f_num_cells * 4));
	FG_CUDA(cudaMemsetAsync(f->cell_color, 0, 3 * (size_t)f->d.f_num_cells * 4, f->st));
	return FLIPGPU_OK;
}

// ---- frame pipeline (see FlipSim).  next_staging: the buffer the next frame goes into (allocated on first use; the
// simulation stream is ordered behind the PCIe copy that may still be reading it -- two frames ago).
int next_staging(FlipSim* f, int* b_out) {
	const int b = 1 - f->frame_idx;
	if (!f->staging[b]) FG_CUDA(cudaMalloc(&f->staging[b], (size_t)std::max(f->d.max_particles, 1) * sizeof(float2)));
	if (f->copy_pending[b]) FG_CUDA(cudaStreamWaitEvent(f->st, f->ev_copied[b], 0));
	*b_out = b;
	return FLIPGPU_OK;
}
int staged(FlipSim* f, int b) {
	FG_CUDA(cudaEventRecord(f->ev_staged[b], f->st));
	f->frame_idx = b;
	f->staged_valid = true;
	return FLIPGPU_OK;
}
// stand-alone un-permute (a frame is asked for that the last step did not stage)
int stage_positions(FlipSim* f) {
	const int np = f->d.num_particles;
	if (np == 0) return FLIPGPU_OK;
	int b = 0;
	FG_TRY(next_staging(f, &b));
	if (f->identity) FG_CUDA(cudaMemcpyAsync(f->staging[b], f->pos[f->cur], (size_t)np * sizeof(float2), cudaMemcpyDeviceToDevice, f->st));
	else { k_scatter_to_caller<float2, 1><<<pgrid(f), kThreads, 0, f->st>>>(np, f->orig[f->cur], f->pos[f->cur], f->staging[b]); f->lc.n++; }
	FG_CUDA(cudaGetLastError());
	return staged(f, b);
}
// Call before anything changes positions / slot order: the staged frame no longer matches the state afterwards.
int positions_will_change(FlipSim* f) {
	f->staged_valid = false;
	return FLIPGPU_OK;
}

// order the simulation stream behind the uploads queued on the upload stream
int join_uploads(FlipSim* f) {
	if (f->upload_pending) FG_CUDA(cudaStreamWaitEvent(f->st, f->ev_uploaded, 0));
	f->upload_pending = false;
	return FLIPGPU_OK;
}

// put every particle array back into caller order (slot k == caller particle k)
int unsort(FlipSim* f) {
	if (f->identity) return FLIPGPU_OK;
	FG_TRY(positions_will_change(f));
	int np = f->d.num_particles, c = f->cur, n = 1 - c;
	if (np > 0) {
		unsigned gr = pgrid(f);
		k_scatter_to_caller<float2, 1><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->pos[c], f->pos[n]);
		k_scatter_to_caller<float2, 1><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->vel[c], f->vel[n]);
		f->lc.n += 2;
		if (f->col[0]) {
			k_scatter_to_caller<float, 3><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->col[c], f->col[n]);
			f->lc.n++;
		}
		k_iota<<<gr, kThreads, 0, f->st>>>(np, f->orig[n]);
		f->lc.n++;
		FG_CUDA(cudaGetLastError());
	}
	f->cur = n;
	f->identity = true;
	f->bins_valid = false;
	return FLIPGPU_OK;
}

// Phase boundary: CUDA event for the per-phase timers (flip_enable_timings) and an NVTX range for nsys / ncu timelines.
const char* const kPhaseNames[FLIP_PH_COUNT + 1] = {"flip:integrate_bin", "flip:separate", "flip:collide_mark", "flip:p2g", "flip:solve", "flip:g2p", "flip:colors", nullptr};
void phase_mark(FlipSim* f, int idx) {
	if (f->timings_on) cudaEventRecord(f->ev[f->ev_slot][idx], f->st);
	if (idx > 0) nvtxRangePop();
	if (kPhaseNames[idx]) nvtxRangePushA(kPhaseNames[idx]);
}
void collect_slot(FlipSim* f, int slot) {
	if (!f->ev_pending[slot]) return;
	cudaEventSynchronize(f->ev[slot][FLIP_PH_COUNT]);
	for (int i = 0; i < FLIP_PH_COUNT; i++) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, f->ev[slot][i], f->ev[slot][i + 1]) == cudaSuccess) f->ph_ms[i] += ms;
	}
	f->ev_pending[slot] = false;
}
void collect_timings(FlipSim* f) {
	for (int s = 0; s < FlipSim::kEvRing; s++) collect_slot(f, s);
}

inline unsigned crowd_grid() { return (unsigned)num_sms() * 4; }
// inclusive running sum of count[0..n) into first[0..n), first[n] = total (flip_fluid.h:180-186)
void running_sum(FlipSim* f, const int* count, int n, int* first) {
	int tiles = (int)cdiv((size_t)n, kScanTile);
	k_scan_tile_sums<<<tiles, kScanThreads, 0, f->st>>>(count, n, f->tile_sum);
	k_scan_tile_offsets<<<1, 1024, 0, f->st>>>(f->tile_sum, tiles);
	k_scan_tiles<<<tiles, kScanThreads, 0, f->st>>>(count, n, f->tile_sum, first);
	f->lc.n += 3;
}

// F2: counting sort into hash-cell order.  keys/count must already hold this step's histogram.
// n_alive >= 0 (sharded steps): entries with kDeadKey are not placed; the reorder covers the n_alive placed ones, which is the
// particle count from here on.
int bin_from_histogram(FlipSim* f, bool integrate = false, float dt = 0.f, float gravity = 0.f, int n_alive = -1) {
	Geo g = geo(f);
	const size_t hoff = (size_t)g.pnx * g.hy0;
	int Hc = g.pnx * (g.hy1 - g.hy0), np = g.np;
	running_sum(f, f->count + hoff, Hc, f->first + hoff);
	int c = f->cur, n = 1 - c;
	if (np > 0) {
		k_fill<<<pgrid(f), kThreads, 0, f->st>>>(np, f->keys, f->orig[c], f->first, f->src, f->orig[n]);
		const CrowdQueue crowd{f->crowd_list, f->crowd_list + kCrowdCap};
		FG_CUDA(cudaMemsetAsync(crowd.count, 0, 4, f->st));
		k_fixup_cells<<<wave_grid(Hc, kThreads), kThreads, 0, f->st>>>(Hc, f->count + hoff, f->first + hoff, f->src, f->orig[n], crowd);
		k_fixup_crowded<true><<<crowd_grid(), 256, 0, f->st>>>(crowd, f->count + hoff, f->first + hoff, f->src, f->orig[n]);
		f->lc.n++;
		Geo gr = g;
		if (n_alive >= 0) gr.np = n_alive;
		const unsigned rgrid = wave_grid((size_t)std::max(gr.np, 1), kThreads);
		if (gr.np > 0) {
			if (f->col[0]) {
				if (integrate) k_reorder<true, true><<<rgrid, kThreads, 0, f->st>>>(gr, f->src, f->pos[c], f->vel[c], f->col[c], f->pos[n], f->vel[n], f->col[n], dt, gravity);
				else k_reorder<true><<<rgrid, kThreads, 0, f->st>>>(gr, f->src, f->pos[c], f->vel[c], f->col[c], f->pos[n], f->vel[n], f->col[n], 0.f, 0.f);
			} else {
				if (integrate) k_reorder<false, true><<<rgrid, kThreads, 0, f->st>>>(gr, f->src, f->pos[c], f->vel[c], nullptr, f->pos[n], f->vel[n], nullptr, dt, gravity);
				else k_reorder<false><<<rgrid, kThreads, 0, f->st>>>(gr, f->src, f->pos[c], f->vel[c], nullptr, f->pos[n], f->vel[n], nullptr, 0.f, 0.f);
			}
		}
		f->lc.n += 3;
		f->cur = n;
		f->identity = false;
	}
	FG_CUDA(cudaGetLastError());
	if (n_alive >= 0) f->d.num_particles = n_alive;
	f->bins_valid = true;
	return FLIPGPU_OK;
}

// F1 + F2 of a whole step: the integration is carried out by the reorder pass of the sort (see k_integrate_bin STORE)
int integrate_sorted(FlipSim* f, float dt, float gravity) {
	Geo g = geo(f);
	FG_TRY(positions_will_change(f));
	int c = f->cur;
	FG_CUDA(cudaMemsetAsync(f->count + (size_t)g.pnx * g.hy0, 0, (size_t)g.pnx * (g.hy1 - g.hy0) * 4, f->st));
	if (g.np > 0) {
		k_integrate_bin<true, true, false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], dt, gravity, f->keys, f->count);
		f->lc.n++;
		FG_CUDA(cudaGetLastError());
	}
	return bin_from_histogram(f, true, dt, gravity);
}
int integrate_and_bin(FlipSim* f, bool integrate, bool bin, float dt, float gravity) {
	Geo g = geo(f);
	FG_TRY(positions_will_change(f));
	int c = f->cur;
	if (bin) FG_CUDA(cudaMemsetAsync(f->count + (size_t)g.pnx * g.hy0, 0, (size_t)g.pnx * (g.hy1 - g.hy0) * 4, f->st));
	if (g.np > 0) {
		if (integrate && bin) k_integrate_bin<true, true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], dt, gravity, f->keys, f->count);
		else if (integrate) k_integrate_bin<true, false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], dt, gravity, f->keys, f->count);
		else if (bin) k_integrate_bin<false, true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], dt, gravity, f->keys, f->count);
		f->lc.n++;
		FG_CUDA(cudaGetLastError());
	}
	if (bin) FG_TRY(bin_from_histogram(f));
	return FLIPGPU_OK;
}


// serial-order separation (exact mode): level schedule from the bins, then the levels one after the other
int separate_exact(FlipSim* f, int iters, bool colors) {
	Geo g = geo(f);
	if (g.np == 0 || iters <= 0) return FLIPGPU_OK;
	if (colors) FG_TRY(ensure_colors(f));
	const int c = f->cur;
	const size_t M = (size_t)std::max(f->d.max_particles, 1);
	if (!f->lvl[0]) {
		FG_CUDA(cudaMalloc(&f->lvl[0], M * 4));
		FG_CUDA(cudaMalloc(&f->lvl[1], M * 4));
		FG_CUDA(cudaMalloc(&f->d_flags, 4 * sizeof(int)));
	}
	// positions at binning time = the current ones (separation follows the binning directly); keep a copy for the schedule
	float2* binpos = f->pos[1 - c];
	FG_CUDA(cudaMemcpyAsync(binpos, f->pos[c], (size_t)g.np * sizeof(float2), cudaMemcpyDeviceToDevice, f->st));
	FG_CUDA(cudaMemsetAsync(f->lvl[0], 0, (size_t)g.np * 4, f->st));
	int a = 0, levels = 0;
	for (int round = 0; round < g.np + 2; round++) {  // Jacobi relaxation of the longest-path problem: converges in (depth) rounds
		FG_CUDA(cudaMemsetAsync(f->d_flags, 0, sizeof(int), f->st));
		k_sep_levels<<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, f->orig[c], binpos, f->lvl[a], f->lvl[1 - a], f->d_flags);
		f->lc.n++;
		a = 1 - a;
		int changed = 0;
		if ((round & 7) == 7 || round < 2) {  // poll the flag every few rounds
			FG_CUDA(cudaMemcpyAsync(&changed, f->d_flags, sizeof(int), cudaMemcpyDeviceToHost, f->st));
			FG_CUDA(cudaStreamSynchronize(f->st));
			if (!changed && round >= 1) {
				// one more quiet round proves the fixed point (the flag only covers the last round)
				break;
			}
		}
	}
	{	// number of levels = max level
		std::vector<int> h(g.np);
		FG_CUDA(cudaMemcpyAsync(h.data(), f->lvl[a], (size_t)g.np * 4, cudaMemcpyDeviceToHost, f->st));
		FG_CUDA(cudaStreamSynchronize(f->st));
		for (int v : h) levels = std::max(levels, v);
	}
	f->sep_levels = levels;
	FG_CUDA(cudaMemsetAsync(f->d_flags + 1, 0, sizeof(int), f->st));
	for (int it = 0; it < iters; it++)
		for (int L = 1; L <= levels; L++) {
			if (colors) k_sep_exact<true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, f->lvl[a], L, f->pos[c], f->col[c], binpos, f->d_flags + 1);
			else k_sep_exact<false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, f->lvl[a], L, f->pos[c], nullptr, binpos, f->d_flags + 1);
			f->lc.n++;
		}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

struct FuseCollide {   // obstacle of the collision pass that the last separation sweep runs itself (flip_step fast path)
	float ox, oy, ovx, ovy, orad;
};
int mark_prepare(FlipSim* f);

int separate(FlipSim* f, int iters, bool colors, const FuseCollide* fuse = nullptr) {
	Geo g = geo(f);
	if (g.np == 0) return FLIPGPU_OK;
	FG_TRY(positions_will_change(f));
	if (f->exact_order) return separate_exact(f, iters, colors);
	if (colors) FG_TRY(ensure_colors(f));
	if (fuse) FG_TRY(mark_prepare(f));   // cell types from s and a cleared list histogram must exist before the fused sweep marks cells
	// Sweeps ping-pong between the two position buffers; velocities (and ids) stay in buffer `cur`,
	// so after an odd number of sweeps the result is copied back rather than swapping every array.
	int c = f->cur;
	float2* a = f->pos[c];
	float2* b = f->pos[1 - c];
	float* ca = colors ? f->col[c] : nullptr;
	float* cb = colors ? f->col[1 - c] : nullptr;
	for (int it = 0; it < iters; it++) {
		if (fuse && it == iters - 1) {
			if (colors) k_separate<true, true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, a, b, ca, cb, f->vel[c], fuse->ox, fuse->oy, fuse->ovx, fuse->ovy, fuse->orad, f->cell_type, f->keys, f->gcount);
			else k_separate<false, true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, a, b, nullptr, nullptr, f->vel[c], fuse->ox, fuse->oy, fuse->ovx, fuse->ovy, fuse->orad, f->cell_type, f->keys, f->gcount);
		} else if (colors) k_separate<true, false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, a, b, ca, cb, nullptr, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr);
		else k_separate<false, false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->first, a, b, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr);
		f->lc.n++;
		std::swap(a, b);
		std::swap(ca, cb);
	}
	FG_CUDA(cudaGetLastError());
	if (a != f->pos[c]) {
		FG_CUDA(cudaMemcpyAsync(f->pos[c], a, (size_t)g.np * sizeof(float2), cudaMemcpyDeviceToDevice, f->st));
		if (colors) FG_CUDA(cudaMemcpyAsync(f->col[c], ca, 3 * (size_t)g.np * 4, cudaMemcpyDeviceToDevice, f->st));
	}
	return FLIPGPU_OK;
}

// cell_type from s (flip_fluid.h:348-350) and an empty histogram for the P2G lists
int mark_prepare(FlipSim* f) {
	Geo g = geo(f);
	FG_TRY(join_uploads(f));   // s (cell types), u, v (solid restore of the P2G) are read from here on
	const size_t goff = (size_t)g.nx * g.gy0;
	const int C = g.nx * (g.gy1 - g.gy0);
	k_cell_type_from_s<<<wave_grid(C, kThreads), kThreads, 0, f->st>>>(C, f->s + goff, f->cell_type + goff);
	f->lc.n++;
	FG_CUDA(cudaMemsetAsync(f->gcount + goff, 0, (size_t)C * 4, f->st));
	return FLIPGPU_OK;
}

// particles_done: the last separation sweep already ran the collision pass, the marking and the list histogram
int collide_mark(FlipSim* f, bool collide, bool mark, float ox, float oy, float ovx, float ovy, float orad, bool particles_done = false) {
	Geo g = geo(f);
	FG_TRY(join_uploads(f));
	if (collide && !particles_done) FG_TRY(positions_will_change(f));
	const size_t goff = (size_t)g.nx * g.gy0;
	int c = f->cur, C = g.nx * (g.gy1 - g.gy0);
	if (mark && !particles_done) FG_TRY(mark_prepare(f));
	if (g.np > 0 && !particles_done) {
		int* ct = mark ? f->cell_type : nullptr;
		if (collide) k_collide_mark<true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], ox, oy, ovx, ovy, orad, ct, f->keys, f->gcount);
		else k_collide_mark<false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], ox, oy, ovx, ovy, orad, ct, f->keys, f->gcount);
		f->lc.n++;
	}
	if (mark) {  // index-only counting sort by stencil base cell: gidx (= src) lists the slots of every grid cell
		running_sum(f, f->gcount + goff, C, f->gfirst + goff);
		if (g.np > 0) {
			k_fill_grid<<<pgrid(f), kThreads, 0, f->st>>>(g.np, f->keys, f->gfirst, f->src);
			if (f->exact_order) k_fixup_grid_by_id<<<wave_grid(C, kThreads), kThreads, 0, f->st>>>(C, f->gcount + goff, f->gfirst + goff, f->src, f->orig[c]);
			else {
				const CrowdQueue crowd{f->crowd_list, f->crowd_list + kCrowdCap};
				FG_CUDA(cudaMemsetAsync(crowd.count, 0, 4, f->st));
				k_fixup_grid<<<wave_grid(C, kThreads), kThreads, 0, f->st>>>(C, f->gcount + goff, f->gfirst + goff, f->src, crowd);
				k_fixup_crowded<false><<<crowd_grid(), 256, 0, f->st>>>(crowd, f->gcount + goff, f->gfirst + goff, f->src, nullptr);
				f->lc.n++;
			}
			f->lc.n += 2;
		}
	}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

int shard_rest_latch(FlipSim* f);

int p2g(FlipSim* f) {
	Geo g = geo(f);
	int c = f->cur;
	dim3 block(32, 8);
	dim3 grid(cdiv(g.nx, block.x), cdiv(g.ny1 - g.ny0, block.y));
	if (f->exact_order)
		k_p2g_gather_exact<<<grid, block, 0, f->st>>>(g, f->gfirst, f->src, f->orig[c], f->pos[c], f->vel[c], f->cell_type, f->u, f->v,
		                                              f->prev_u, f->prev_v, f->dens, f->p);
	else {
		// FLIPGPU_P2G_V1=1 runs the general loop at every node (A/B timing; results are identical)
		const char* v1 = getenv("FLIPGPU_P2G_V1");
		if (v1 && atoi(v1) != 0)
			k_p2g_gather<false><<<grid, block, 0, f->st>>>(g, f->gfirst, f->src, f->pos[c], f->vel[c], f->cell_type, f->u, f->v,
			                                               f->prev_u, f->prev_v, f->dens, f->p);
		else
			k_p2g_gather<true><<<grid, block, 0, f->st>>>(g, f->gfirst, f->src, f->pos[c], f->vel[c], f->cell_type, f->u, f->v,
			                                              f->prev_u, f->prev_v, f->dens, f->p);
	}
	f->lc.n++;
	if (!f->rest_seen && f->sharded) {
		FG_TRY(shard_rest_latch(f));
	} else if (!f->rest_seen && f->exact_order) {
		k_rest_latch_serial<<<1, 32, 0, f->st>>>(f->d.f_num_cells, f->cell_type, f->dens, f->d_rest);
		f->lc.n++;
		k_word_to_host<<<1, 32, 0, f->st>>>((const int*)f->d_rest, (volatile int*)f->h_pinned);
	} else if (!f->rest_seen) {
		k_rest_partials<<<kPartials, 256, 0, f->st>>>(f->d.f_num_cells, f->cell_type, f->dens, f->d_rest, f->d_partials);
		k_rest_latch<<<1, 32, 0, f->st>>>(kPartials, f->d_partials, f->d_rest);
		f->lc.n += 2;
		// mirror the scalar into pinned memory; once a later call sees it non-zero the latch kernels stop
		k_word_to_host<<<1, 32, 0, f->st>>>((const int*)f->d_rest, (volatile int*)f->h_pinned);
	}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

SorGrid sor_view(FlipSim* f) {
	SorGrid g;
	g.nf = f->d.f_num_x; g.ns = f->d.f_num_y;
	g.vel_f = f->u; g.vel_s = f->v; g.p = f->p; g.s = f->s; g.cell_type = f->cell_type;
	g.density = f->dens; g.rest = f->d_rest; g.x_fast = true;
	g.alt_f = f->u_alt; g.alt_s = f->v_alt; g.alt_p = f->p_alt;
	return g;
}
int ensure_alt(FlipSim* f) {
	if (f->u_alt) return FLIPGPU_OK;
	size_t C = f->d.f_num_cells;
	FG_CUDA(cudaMalloc(&f->u_alt, C * 4)); FG_CUDA(cudaMalloc(&f->v_alt, C * 4)); FG_CUDA(cudaMalloc(&f->p_alt, C * 4));
	FG_CUDA(cudaMemsetAsync(f->u_alt, 0, C * 4, f->st)); FG_CUDA(cudaMemsetAsync(f->v_alt, 0, C * 4, f->st)); FG_CUDA(cudaMemsetAsync(f->p_alt, 0, C * 4, f->st));
	return FLIPGPU_OK;
}
// s was written by the caller: check on the device (asynchronously) whether it is still 0/1-valued
int queue_s_check(FlipSim* f, size_t first_cell, size_t n_cells, bool whole_array, cudaStream_t st) {
	if (n_cells == 0) return FLIPGPU_OK;
	// a whole-array upload starts from scratch; a partial one adds to what is known (or still being checked) about the rest
	if (whole_array) FG_CUDA(cudaMemsetAsync(f->d_sflag, 0, sizeof(int), st));
	else if (!f->s_check_pending) FG_CUDA(cudaMemsetAsync(f->d_sflag, f->s_binary ? 0 : 0xff, sizeof(int), st));
	k_check_s_binary<<<wave_grid(n_cells, kThreads), kThreads, 0, st>>>(n_cells, f->s + first_cell, f->d_sflag);
	f->lc.n++;
	FG_CUDA(cudaGetLastError());
	k_word_to_host<<<1, 32, 0, st>>>(f->d_sflag, (volatile int*)f->h_sflag);
	FG_CUDA(cudaEventRecord(f->ev_scheck, st));
	f->s_check_pending = true;
	return FLIPGPU_OK;
}
// ... and wait for the answer only where the solver is chosen (the upload and the check are long done by then)
int resolve_s_binary(FlipSim* f) {
	if (!f->s_check_pending) return FLIPGPU_OK;
	FG_CUDA(cudaEventSynchronize(f->ev_scheck));
	f->s_binary = *f->h_sflag == 0;
	f->s_check_pending = false;
	return FLIPGPU_OK;
}

// the projection of one step; an odd number of ping-pong rounds leaves the result in the partner arrays -> swap the names
int solve(FlipSim* f, int iters, float cp, float omega, bool drift, int order) {
	FG_TRY(resolve_s_binary(f));
	if (order == FLIP_SOLVER_RED_BLACK && f->s_binary) FG_TRY(ensure_alt(f));
	bool swapped = false;
	FG_TRY(sor_solve(&f->ws, sor_view(f), iters, cp, omega, drift, order, f->s_binary, f->st, &f->lc, &swapped));
	if (swapped) { std::swap(f->u, f->u_alt); std::swap(f->v, f->v_alt); std::swap(f->p, f->p_alt); }
	return FLIPGPU_OK;
}

int g2p(FlipSim* f, float flip_ratio, bool stage = false) {
	Geo g = geo(f);
	if (g.np == 0) return FLIPGPU_OK;
	int c = f->cur;
	if (stage && !f->identity) {
		int b = 0;
		FG_TRY(next_staging(f, &b));
		k_g2p<true><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], f->u, f->v, f->prev_u, f->prev_v, f->cell_type, flip_ratio, f->orig[c], f->staging[b], nullptr);
		f->lc.n++;
		FG_CUDA(cudaGetLastError());
		return staged(f, b);
	}
	k_g2p<false><<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[c], f->vel[c], f->u, f->v, f->prev_u, f->prev_v, f->cell_type, flip_ratio, nullptr, nullptr, f->sharded ? f->d_counters + 8 : nullptr);
	f->vmax_valid = f->sharded;
	f->lc.n++;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}


// ------------------------------------------------------------------ slab-sharded step
int shard_rest_latch(FlipSim* f) {
	const size_t off = (size_t)f->d.f_num_x * f->j0;
	const int n = f->d.f_num_x * (f->j1 - f->j0);
	k_rest_partials<<<kPartials, 256, 0, f->st>>>(n, f->cell_type + off, f->dens + off, f->d_rest, f->d_partials);
	float* pair = f->d_partials + 2 * kPartials;
	k_rest_pair<<<1, 32, 0, f->st>>>(kPartials, f->d_partials, f->d_rest, pair);
	if (f->nranks > 1) FG_NCCL(g_nccl.AllReduce(pair, pair, 2, ncclFloat, ncclSum, f->comm, f->st));
	k_rest_from_pair<<<1, 32, 0, f->st>>>(pair, f->d_rest);
	f->lc.n += 3;
	k_word_to_host<<<1, 32, 0, f->st>>>((const int*)f->d_rest, (volatile int*)f->h_pinned);
	return FLIPGPU_OK;
}

// swap `rows` rows at both slab boundaries: my outermost own rows go into the neighbour's halo and vice versa.
// Arrays keep the global size, so a row sits at the same offset on every rank.  (One send/recv pair per field: packing the
// fields into one message per neighbour was measured and changed nothing -- 12.9 vs 13.0 ms per step at 2 GPUs; what the
// phase table shows around an exchange is the wait for the slower rank, not the transfer.)
int shard_exchange_rows(FlipSim* f, void* const* fields, int n_fields, int rows) {
	if (f->nranks == 1) return FLIPGPU_OK;
	const size_t nx = f->d.f_num_x, blk = nx * rows;
	const bool has_up = f->rank + 1 < f->nranks, has_down = f->rank > 0;
	FG_NCCL(g_nccl.GroupStart());
	for (int k = 0; k < n_fields; k++) {
		float* a = (float*)fields[k];
		if (has_up) {
			FG_NCCL(g_nccl.Send(a + nx * (f->j1 - rows), blk, ncclFloat, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(a + nx * f->j1, blk, ncclFloat, f->rank + 1, f->comm, f->st));
		}
		if (has_down) {
			FG_NCCL(g_nccl.Send(a + nx * f->j0, blk, ncclFloat, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(a + nx * (f->j0 - rows), blk, ncclFloat, f->rank - 1, f->comm, f->st));
		}
	}
	FG_NCCL(g_nccl.GroupEnd());
	f->exchanges++;
	return FLIPGPU_OK;
}

// Stable compaction of buffer `cur` into buffer `1 - cur`: the owned particles, and with `exchange` the boundary copies for both
// neighbours into the send buffers.  Leaves the three counts in d_counters[0..2].
int shard_compact(FlipSim* f, bool exchange) {
	Geo g = geo(f);
	if (g.np <= 0) return FLIPGPU_OK;   // (the counters were cleared by the caller)
	const int c = f->cur, n = 1 - c;
	const bool has_up = exchange && f->rank + 1 < f->nranks, has_down = exchange && f->rank > 0;
	const int n_chunks = (int)cdiv((size_t)g.np, (size_t)kShardChunk);
	const int* band = exchange ? f->d_counters + 6 : nullptr;
	ShardPack keep{f->pos[n], f->vel[n], f->orig[n]}, down{f->send_pos[0], f->send_vel[0], f->send_id[0]}, up{f->send_pos[1], f->send_vel[1], f->send_id[1]};
	const unsigned grid = wave_grid((size_t)n_chunks * 32, kThreads);
	int* incl = f->d_chunk + f->chunk_stride;
	k_shard_count<<<grid, kThreads, 0, f->st>>>(g, f->j0, f->j1, band, has_down, has_up, f->pos[c], n_chunks, f->d_chunk);
	running_sum(f, f->d_chunk, 3 * n_chunks, incl);
	k_shard_place<<<grid, kThreads, 0, f->st>>>(g, f->j0, f->j1, band, has_down, has_up, f->pos[c], f->vel[c], f->orig[c], keep, down, up, n_chunks, f->d_chunk,
	                                            incl, f->cap_send, f->d_counters);
	f->lc.n += 2;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

// keep what this rank owns (by position), refresh the ghosts from both neighbours
int shard_exchange_particles(FlipSim* f, float dt, float gravity, int sep_iters, int* n_alive) {
	Geo g = geo(f);
	const int c = f->cur;
	const bool has_up = f->rank + 1 < f->nranks, has_down = f->rank > 0;
	FG_TRY(positions_will_change(f));
	FG_CUDA(cudaMemsetAsync(f->d_counters, 0, 8 * sizeof(int), f->st));
	FG_CUDA(cudaMemsetAsync(f->count + (size_t)g.pnx * g.hy0, 0, (size_t)g.pnx * (g.hy1 - g.hy0) * 4, f->st));   // this step's hash histogram
	// ghost band for this step from the fastest particle anywhere (device-side: no host round trip).  The G2P of the last
	// step left the maximum behind; after an upload (or before the first step) it is taken here.
	if (!f->vmax_valid) {
		FG_CUDA(cudaMemsetAsync(f->d_counters + 8, 0, sizeof(int), f->st));
		if (g.np > 0) { k_shard_vmax<<<pgrid(f), kThreads, 0, f->st>>>(g.np, f->vel[c], f->d_counters + 8); f->lc.n++; }
	}
	f->vmax_valid = false;
	if (f->nranks > 1) FG_NCCL(g_nccl.AllReduce(f->d_counters + 8, f->d_counters + 8, 1, ncclInt, ncclMax, f->comm, f->st));
	k_shard_band<<<1, 32, 0, f->st>>>(f->d_counters + 8, dt, gravity, f->d.h, f->d.particle_radius, sep_iters, f->d_counters + 6);
	f->lc.n++;
	const int held = g.np;   // owned + last step's ghosts, in hash-cell order in buffer `cur`; arrivals are appended behind them
	if (held > 0) {
		ShardPack down{f->send_pos[0], f->send_vel[0], f->send_id[0]}, up{f->send_pos[1], f->send_vel[1], f->send_id[1]};
		k_shard_key<<<pgrid(f), kThreads, 0, f->st>>>(g, f->j0, f->j1, f->d_counters + 6, has_down, has_up, f->pos[c], f->vel[c], f->orig[c], dt, gravity,
		                                              f->keys, f->count, down, up, f->cap_send, f->d_counters);
		f->lc.n++;
	}
	// tell each neighbour how many particles are coming (device-to-device ints), then read the counters once
	if (f->nranks > 1) {
		FG_NCCL(g_nccl.GroupStart());
		if (has_down) {
			FG_NCCL(g_nccl.Send(f->d_counters + 1, 1, ncclInt, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->d_counters + 3, 1, ncclInt, f->rank - 1, f->comm, f->st));
		}
		if (has_up) {
			FG_NCCL(g_nccl.Send(f->d_counters + 2, 1, ncclInt, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->d_counters + 4, 1, ncclInt, f->rank + 1, f->comm, f->st));
		}
		FG_NCCL(g_nccl.GroupEnd());
	}
	k_shard_check<<<1, 32, 0, f->st>>>(f->d_counters, f->cap_send, held, f->d.max_particles);
	f->lc.n++;
	if (f->nranks > 1) FG_NCCL(g_nccl.AllReduce(f->d_counters + 7, f->d_counters + 7, 1, ncclInt, ncclMax, f->comm, f->st));
	FG_CUDA(cudaMemcpyAsync(f->h_counters, f->d_counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, f->st));
	FG_CUDA(cudaStreamSynchronize(f->st));
	const int n_keep = f->h_counters[0], n_down = f->h_counters[1], n_up = f->h_counters[2], from_down = f->h_counters[3], from_up = f->h_counters[4];
	FG_CHECK(f->h_counters[7] == 0, FLIPGPU_ESTATE, "shard: a rank ran out of room in this step's particle exchange (here: %d held + %d + %d received of %d slots, "
	         "%d / %d boundary particles of %d send slots) -- every rank stops", held, from_down, from_up, f->d.max_particles, n_down, n_up, f->cap_send);
	f->band = f->h_counters[6];
	// (the band is the same number on every rank and min_slab_rows is too: all ranks take this exit together, none is left
	// blocked in the NCCL group below)
	FG_CHECK(f->nranks == 1 || f->band <= f->min_slab_rows - 1, FLIPGPU_ESTATE,
	         "shard: the ghost band (%d rows: fastest particle, %d separation sweeps) is wider than the thinnest slab (%d rows) -- use fewer ranks or a smaller dt",
	         f->band, sep_iters, f->min_slab_rows);
	FG_CHECK(n_down <= f->cap_send && n_up <= f->cap_send, FLIPGPU_ESTATE, "shard: %d/%d boundary particles exceed the send capacity %d", n_down, n_up, f->cap_send);
	FG_CHECK((long long)held + from_down + from_up <= f->d.max_particles, FLIPGPU_ESTATE, "shard: %d held + %d + %d received exceed max_particles %d",
	         held, from_down, from_up, f->d.max_particles);
	if (f->nranks > 1) {
		FG_NCCL(g_nccl.GroupStart());
		int at = held;
		if (has_down) {
			FG_NCCL(g_nccl.Send(f->send_pos[0], 2 * (size_t)n_down, ncclFloat, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Send(f->send_vel[0], 2 * (size_t)n_down, ncclFloat, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Send(f->send_id[0], (size_t)n_down, ncclInt, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->pos[c] + at, 2 * (size_t)from_down, ncclFloat, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->vel[c] + at, 2 * (size_t)from_down, ncclFloat, f->rank - 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->orig[c] + at, (size_t)from_down, ncclInt, f->rank - 1, f->comm, f->st));
			at += from_down;
		}
		if (has_up) {
			FG_NCCL(g_nccl.Send(f->send_pos[1], 2 * (size_t)n_up, ncclFloat, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Send(f->send_vel[1], 2 * (size_t)n_up, ncclFloat, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Send(f->send_id[1], (size_t)n_up, ncclInt, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->pos[c] + at, 2 * (size_t)from_up, ncclFloat, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->vel[c] + at, 2 * (size_t)from_up, ncclFloat, f->rank + 1, f->comm, f->st));
			FG_NCCL(g_nccl.Recv(f->orig[c] + at, (size_t)from_up, ncclInt, f->rank + 1, f->comm, f->st));
		}
		FG_NCCL(g_nccl.GroupEnd());
		f->exchanges++;
	}
	const int n_recv = from_down + from_up;
	if (n_recv > 0) {  // the arrivals (ghosts, migrants): key + histogram from the integrated position like the owned ones
		Geo gt = g;
		gt.np = n_recv;
		k_integrate_bin<true, true, false><<<wave_grid((size_t)n_recv, kThreads), kThreads, 0, f->st>>>(gt, f->pos[c] + held, f->vel[c] + held, dt, gravity, f->keys + held, f->count);
		f->lc.n++;
	}
	FG_CUDA(cudaGetLastError());
	f->n_owned = n_keep;
	f->d.num_particles = held + n_recv;   // what the sort reads; it leaves n_keep + n_recv (dead entries drop out)
	f->identity = false;
	f->bins_valid = false;
	*n_alive = n_keep + n_recv;
	return FLIPGPU_OK;
}

// red-black sweeps with the communication-avoiding deep halo of sor_shard.cu, on this handle's (global-size) arrays.
// With 0/1-valued s every exchange is followed by ONE temporally blocked launch that runs the 8 colour passes the 9-row
// halo allows (sor_rb_tiled_kernel); otherwise one launch per colour pass on the shrinking row band.
int shard_solve(FlipSim* f, int iters, float cp, float omega, bool drift) {
	const int H = f->halo, K = std::max(H - 1, 1), ny = f->d.f_num_y;
	const bool has_up = f->rank + 1 < f->nranks, has_down = f->rank > 0;
	FG_TRY(resolve_s_binary(f));  // (both branches below issue the same NCCL exchanges, so ranks may differ in this)
	if (f->s_binary && K == sor_rb_passes_per_round()) {
		FG_TRY(ensure_alt(f));
		FG_TRY(sor_rb_prep(&f->ws, sor_view(f), drift, f->gy0, f->gy1, f->st, &f->lc));
		const int out_lo = has_down ? f->j0 - 1 : f->j0, out_hi = has_up ? f->j1 + 1 : f->j1;   // own rows +- 1: completes the shared face rows
		for (int pass0 = 0; pass0 < 2 * iters; pass0 += K) {
			void* vel[2] = {f->u, f->v};
			FG_TRY(shard_exchange_rows(f, vel, 2, H));
			FG_TRY(sor_rb_round(f->ws, sor_view(f), f->u_alt, f->v_alt, f->p_alt, std::min(K, 2 * iters - pass0), pass0 & 1, std::max(out_lo, 0),
			                    std::min(out_hi, ny), cp, omega, drift, f->st, &f->lc));
			std::swap(f->u, f->u_alt); std::swap(f->v, f->v_alt); std::swap(f->p, f->p_alt);
		}
		return FLIPGPU_OK;
	}
	SorGrid g = sor_view(f);
	void* vel[2] = {f->u, f->v};
	int pass = 0;
	for (int it = 0; it < iters; it++)
		for (int colour = 0; colour < 2; colour++, pass++) {
			const int m = f->nranks > 1 ? pass % K : 0;
			if (f->nranks > 1 && m == 0) FG_TRY(shard_exchange_rows(f, vel, 2, H));
			int lo = has_down ? f->j0 - (H - 1) + m : f->j0, hi = has_up ? f->j1 + (H - 1) - m - 1 : f->j1 - 1;
			lo = std::max(lo, 1); hi = std::min(hi, ny - 2);
			if (hi < lo) continue;
			sor_rb_rows_launch(g, colour, 0, lo, hi, cp, omega, drift && g.density && g.rest, f->st);
			f->lc.n++;
		}
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

int shard_step(FlipSim* f, const FlipStepParams* sp) {
	FG_CHECK(sp->solver_order == FLIP_SOLVER_RED_BLACK, FLIPGPU_EINVAL, "a sharded simulation runs the red-black order (the lexicographic wavefront does not shard)");
	FG_CHECK(!sp->update_colors, FLIPGPU_EINVAL, "render-state colours are not kept in sharded mode");
	const float cp = f->d.density * f->d.h / sp->dt;
	// phase timings (flip_enable_timings): the exchanges are booked on the phase they feed -- particle exchange under
	// integrate_bin, post-P2G halo rows under p2g, post-solve halo rows under g2p
	if (f->timings_on) collect_slot(f, f->ev_slot);
	phase_mark(f, FLIP_PH_INTEGRATE_BIN);
	int n_alive = 0;
	FG_TRY(shard_exchange_particles(f, sp->dt, sp->gravity, sp->separate_particles ? sp->num_particle_iters : 0, &n_alive));    // migrants change owner, ghosts are refreshed; sort keys
	// (the exchange synchronised the stream, so the previous step -- and its copy of the rest density -- is complete on
	// every rank: all ranks take the same branch below and issue the same NCCL calls)
	if (!f->rest_seen && f->h_pinned[0] != 0.f) f->rest_seen = true;
	FG_TRY(bin_from_histogram(f, true, sp->dt, sp->gravity, n_alive));     // owned + ghosts; F1 happens in the reorder pass; what the rank no longer holds drops out
	phase_mark(f, FLIP_PH_SEPARATE);
	// as in the one-GPU step, the last sweep runs the collision / marking pass itself (identical results)
	const bool fused = sp->separate_particles && sp->num_particle_iters >= 1 && f->d.num_particles > 0;
	const FuseCollide fc{sp->obstacle_x, sp->obstacle_y, sp->obstacle_vel_x, sp->obstacle_vel_y, sp->obstacle_radius};
	if (sp->separate_particles) FG_TRY(separate(f, sp->num_particle_iters, false, fused ? &fc : nullptr));
	phase_mark(f, FLIP_PH_COLLIDE_MARK);
	FG_TRY(collide_mark(f, true, true, sp->obstacle_x, sp->obstacle_y, sp->obstacle_vel_x, sp->obstacle_vel_y, sp->obstacle_radius, fused));
	phase_mark(f, FLIP_PH_P2G);
	FG_TRY(p2g(f));                                                        // own node rows
	{	// the rows next to the slab come from their owners: post-P2G fields, cell types, densities, s
		void* fields[] = {f->u, f->v, f->prev_u, f->prev_v, f->dens, f->cell_type, f->s};
		FG_TRY(shard_exchange_rows(f, fields, 7, f->halo));
	}
	phase_mark(f, FLIP_PH_SOLVE);
	FG_TRY(shard_solve(f, sp->num_pressure_iters, cp, sp->over_relaxation, sp->compensate_drift != 0));
	phase_mark(f, FLIP_PH_G2P);
	{	// post-solve velocities (and the pressure) of the halo rows, for the G2P of particles near the boundary
		void* fields[] = {f->u, f->v, f->p};
		FG_TRY(shard_exchange_rows(f, fields, 3, f->halo));
	}
	FG_TRY(g2p(f, sp->flip_ratio));
	phase_mark(f, FLIP_PH_COLORS);
	phase_mark(f, FLIP_PH_COUNT);
	if (f->timings_on) {
		f->ev_pending[f->ev_slot] = true;
		f->ev_slot = (f->ev_slot + 1) % FlipSim::kEvRing;
	}
	return FLIPGPU_OK;
}

int colors(FlipSim* f) {
	Geo g = geo(f);
	FG_TRY(ensure_colors(f));
	FG_TRY(ensure_cell_colors(f));
	if (g.np > 0) {
		k_particle_colors<<<pgrid(f), kThreads, 0, f->st>>>(g, f->pos[f->cur], f->col[f->cur], f->dens, f->d_rest);
		f->lc.n++;
	}
	k_cell_colors<<<wave_grid(f->d.f_num_cells, kThreads), kThreads, 0, f->st>>>(f->d.f_num_cells, f->cell_type, f->dens, f->d_rest, f->cell_color);
	f->lc.n++;
	FG_CUDA(cudaGetLastError());
	return FLIPGPU_OK;
}

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t field_count(const FlipSim* f, int field) {
	size_t C = f->d.f_num_cells, P = f->d.num_particles, H = f->d.p_num_cells;
	switch (field) {
		case FLIP_CELL_COLOR: return 3 * C;
		case FLIP_PARTICLE_POS: case FLIP_PARTICLE_VEL: return 2 * P;
		case FLIP_PARTICLE_COLOR: return 3 * P;
		case FLIP_NUM_CELL_PARTICLES: return H;
		case FLIP_FIRST_CELL_PARTICLE: return H + 1;
		case FLIP_CELL_PARTICLE_IDS: return P;
		default: return C;
	}
}
void* grid_field_ptr(FlipSim* f, int field) {
	switch (field) {
		case FLIP_U: return f->u; case FLIP_V: return f->v;
		case FLIP_DU: case FLIP_DV: case FLIP_PARTICLE_DENSITY: return f->dens;  // one array, see k_p2g_gather
		case FLIP_PREV_U: return f->prev_u; case FLIP_PREV_V: return f->prev_v;
		case FLIP_P: return f->p; case FLIP_S: return f->s; case FLIP_CELL_TYPE: return f->cell_type;
		default: return nullptr;
	}
}
}  // namespace

extern "C" {

const char* flipgpu_last_error(void) { return fg::g_err; }
int flipgpu_version(void) { return 100; }
int flipgpu_device_count(int* count) {
	FG_CHECK(count, FLIPGPU_EINVAL, "count is null");
	cudaError_t e = cudaGetDeviceCount(count);
	if (e != cudaSuccess) { *count = 0; fg::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return FLIPGPU_ENODEV; }
	return FLIPGPU_OK;
}

int flip_create(const FlipParams* pr, FlipSim** out) {
	FG_CHECK(pr && out, FLIPGPU_EINVAL, "flip_create: null argument");
	FG_CHECK(pr->spacing > 0 && pr->particle_radius > 0 && pr->width > 0 && pr->height > 0 && pr->max_particles >= 0,
	         FLIPGPU_EINVAL, "flip_create: non-positive geometry");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		fg::set_error("flip_create: no CUDA device (this library has no CPU fallback)");
		return FLIPGPU_ENODEV;
	}
	int dev = pr->device;
	if (dev < 0) FG_CUDA(cudaGetDevice(&dev));
	FG_CHECK(dev < ndev, FLIPGPU_EINVAL, "flip_create: device %d of %d", dev, ndev);
	FlipSim* f = new FlipSim;
	f->device = dev;
	DeviceGuard guard(dev);
	// derived sizes in fp32, verbatim order of flip_fluid.h:69-74,103-106
	FlipDims& d = f->d;
	d.density = pr->density;
	d.f_num_x = (int)(1 + pr->width / pr->spacing);
	d.f_num_y = (int)(1 + pr->height / pr->spacing);
	d.f_num_cells = d.f_num_x * d.f_num_y;
	d.h = std::max(pr->width / d.f_num_x, pr->height / d.f_num_y);
	d.f_inv_spacing = 1 / d.h;
	d.max_particles = pr->max_particles;
	d.particle_radius = pr->particle_radius;
	d.p_inv_spacing = 1 / (2.2f * pr->particle_radius);
	d.p_num_x = (int)(1 + pr->width * d.p_inv_spacing);
	d.p_num_y = (int)(1 + pr->height * d.p_inv_spacing);
	long long hc = (long long)d.p_num_x * d.p_num_y;
	if (hc >= (1ll << 31) - 8 || (long long)d.f_num_x * d.f_num_y >= (1ll << 31)) {
		delete f;
		fg::set_error("flip_create: grid too large for int32 indices (shard it)");
		return FLIPGPU_EINVAL;
	}
	d.p_num_cells = d.p_num_x * d.p_num_y;
	d.num_particles = 0;
	FG_CUDA(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
	FG_CUDA(cudaStreamCreateWithFlags(&f->st_copy, cudaStreamNonBlocking));
	FG_CUDA(cudaStreamCreateWithFlags(&f->st_up, cudaStreamNonBlocking));
	FG_CUDA(cudaEventCreateWithFlags(&f->ev_fence, cudaEventDisableTiming));
	FG_CUDA(cudaEventCreateWithFlags(&f->ev_uploaded, cudaEventDisableTiming));
	for (int b = 0; b < 2; b++) {
		FG_CUDA(cudaEventCreateWithFlags(&f->ev_staged[b], cudaEventDisableTiming));
		FG_CUDA(cudaEventCreateWithFlags(&f->ev_copied[b], cudaEventDisableTiming));
	}
	size_t C = d.f_num_cells, M = std::max(d.max_particles, 1), H = d.p_num_cells;
	float** grids[] = {&f->u, &f->v, &f->prev_u, &f->prev_v, &f->p, &f->s, &f->dens};
	for (float** gp : grids) {
		FG_CUDA(cudaMalloc(gp, C * 4));
		FG_CUDA(cudaMemsetAsync(*gp, 0, C * 4, f->st));
	}
	FG_CUDA(cudaMalloc(&f->cell_type, C * 4));
	FG_CUDA(cudaMemsetAsync(f->cell_type, 0, C * 4, f->st));
	for (int b = 0; b < 2; b++) {
		FG_CUDA(cudaMalloc(&f->pos[b], M * sizeof(float2)));
		FG_CUDA(cudaMalloc(&f->vel[b], M * sizeof(float2)));
		FG_CUDA(cudaMalloc(&f->orig[b], M * 4));
		FG_CUDA(cudaMemsetAsync(f->pos[b], 0, M * sizeof(float2), f->st));
		FG_CUDA(cudaMemsetAsync(f->vel[b], 0, M * sizeof(float2), f->st));
		k_iota<<<wave_grid(M, kThreads), kThreads, 0, f->st>>>((int)M, f->orig[b]);
	}
	FG_CUDA(cudaMalloc(&f->keys, M * 4));
	FG_CUDA(cudaMalloc(&f->src, M * 4));
	FG_CUDA(cudaMalloc(&f->gcount, C * 4));
	FG_CUDA(cudaMalloc(&f->gfirst, (C + 1) * 4));
	FG_CUDA(cudaMalloc(&f->count, H * 4));
	FG_CUDA(cudaMalloc(&f->first, (H + 1) * 4));
	FG_CUDA(cudaMemsetAsync(f->count, 0, H * 4, f->st));
	FG_CUDA(cudaMemsetAsync(f->first, 0, (H + 1) * 4, f->st));
	FG_CUDA(cudaMemsetAsync(f->src, 0, M * 4, f->st));
	FG_CUDA(cudaMalloc(&f->tile_sum, (size_t)(cdiv(std::max(H, C), kScanTile) + 1) * 4));
	FG_CUDA(cudaMalloc(&f->crowd_list, ((size_t)kCrowdCap + 1) * 4));
	FG_CUDA(cudaMalloc(&f->d_rest, 4));
	FG_CUDA(cudaMemsetAsync(f->d_rest, 0, 4, f->st));
	FG_CUDA(cudaMalloc(&f->d_partials, (3 * kPartials + 8) * 4));
	FG_CUDA(cudaMallocHost(&f->h_pinned, 64));
	f->h_pinned[0] = 0.f;
	FG_CUDA(cudaMalloc(&f->d_sflag, sizeof(int)));
	FG_CUDA(cudaMallocHost(&f->h_sflag, sizeof(int)));
	*f->h_sflag = 0;
	FG_CUDA(cudaEventCreateWithFlags(&f->ev_scheck, cudaEventDisableTiming));
	for (int r = 0; r < FlipSim::kEvRing; r++)
		for (int i = 0; i <= FLIP_PH_COUNT; i++) FG_CUDA(cudaEventCreate(&f->ev[r][i]));
	FG_CUDA(cudaStreamSynchronize(f->st));
	*out = f;
	return FLIPGPU_OK;
}

int flip_destroy(FlipSim* f) {
	if (!f) return FLIPGPU_OK;
	DeviceGuard guard(f->device);
	cudaStreamSynchronize(f->st);
	void* ptrs[] = {f->u, f->v, f->prev_u, f->prev_v, f->p, f->s, f->dens, f->cell_type, f->cell_color, f->pos[0], f->pos[1],
	                f->vel[0], f->vel[1], f->col[0], f->col[1], f->orig[0], f->orig[1], f->keys, f->src, f->gcount, f->gfirst, f->count, f->u_alt, f->v_alt, f->p_alt,
	                f->first, f->tile_sum, f->crowd_list, f->d_rest, f->d_partials};
	for (void* p : ptrs) if (p) cudaFree(p);
	for (auto& G : f->graphs) if (G.exec) cudaGraphExecDestroy(G.exec);
	if (f->comm) g_nccl.CommDestroy(f->comm);
	for (int k = 0; k < 2; k++) { cudaFree(f->send_pos[k]); cudaFree(f->send_vel[k]); cudaFree(f->send_id[k]); }
	cudaFree(f->d_chunk);
	cudaFree(f->d_counters);
	cudaFree(f->lvl[0]); cudaFree(f->lvl[1]); cudaFree(f->d_flags);
	if (f->h_counters) cudaFreeHost(f->h_counters);
	if (f->h_pinned) cudaFreeHost(f->h_pinned);
	cudaFree(f->d_sflag);
	if (f->h_sflag) cudaFreeHost(f->h_sflag);
	if (f->ev_scheck) cudaEventDestroy(f->ev_scheck);
	if (f->st_copy) { cudaStreamSynchronize(f->st_copy); cudaStreamDestroy(f->st_copy); }
	if (f->st_up) { cudaStreamSynchronize(f->st_up); cudaStreamDestroy(f->st_up); }
	if (f->ev_fence) cudaEventDestroy(f->ev_fence);
	if (f->ev_uploaded) cudaEventDestroy(f->ev_uploaded);
	for (int b = 0; b < 2; b++) {
		if (f->ev_staged[b]) cudaEventDestroy(f->ev_staged[b]);
		if (f->ev_copied[b]) cudaEventDestroy(f->ev_copied[b]);
		if (f->staging[b]) cudaFree(f->staging[b]);
	}
	sor_workspace_free(f->ws);
	for (int r = 0; r < FlipSim::kEvRing; r++)
		for (int i = 0; i <= FLIP_PH_COUNT; i++) if (f->ev[r][i]) cudaEventDestroy(f->ev[r][i]);
	cudaStreamDestroy(f->st);
	delete f;
	return FLIPGPU_OK;
}

int flip_get_dims(FlipSim* f, FlipDims* dims) {
	FG_CHECK(f && dims, FLIPGPU_EINVAL, "flip_get_dims: null argument");
	*dims = f->d;
	return FLIPGPU_OK;
}

int flip_set_num_particles(FlipSim* f, int n) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(n >= 0 && n <= f->d.max_particles, FLIPGPU_EINVAL, "num_particles %d outside [0,%d]", n, f->d.max_particles);
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "sharded simulation: the particle count follows flip_upload_particles");
	DeviceGuard guard(f->device);
	FG_TRY(unsort(f));
	FG_TRY(positions_will_change(f));
	f->d.num_particles = n;
	f->bins_valid = false;
	return FLIPGPU_OK;
}

static int upload_impl(FlipSim* f, int field, const void* host, size_t count, bool wait) {
	FG_CHECK(f && host, FLIPGPU_EINVAL, "flip_upload: null argument");
	FG_CHECK(field >= 0 && field < FLIP_FIELD_COUNT, FLIPGPU_EINVAL, "flip_upload: unknown field %d", field);
	DeviceGuard guard(f->device);
	size_t want = field_count(f, field);
	FG_CHECK(count == want, FLIPGPU_EINVAL, "flip_upload: field %d takes %zu elements, got %zu", field, want, count);
	void* dst = grid_field_ptr(f, field);
	FG_CHECK(dst || !f->sharded, FLIPGPU_ESTATE, "sharded simulation: use flip_upload_particles / flip_upload_rows");
	if (!dst) {
		switch (field) {
			case FLIP_PARTICLE_POS: FG_TRY(unsort(f)); FG_TRY(positions_will_change(f)); dst = f->pos[f->cur]; f->bins_valid = false; break;
			case FLIP_PARTICLE_VEL: FG_TRY(unsort(f)); dst = f->vel[f->cur]; break;
			case FLIP_PARTICLE_COLOR: FG_TRY(ensure_colors(f)); FG_TRY(unsort(f)); dst = f->col[f->cur]; break;
			case FLIP_CELL_COLOR: FG_TRY(ensure_cell_colors(f)); dst = f->cell_color; break;
			default:
				fg::set_error("flip_upload: field %d is derived state (hash bins) and cannot be uploaded", field);
				return FLIPGPU_EINVAL;
		}
	}
	if (!wait && !f->sharded && (field == FLIP_S || field == FLIP_U || field == FLIP_V)) {
		// per-frame caller writes: on the upload stream, behind the readers of the old values (everything queued so far)
		FG_CUDA(cudaEventRecord(f->ev_fence, f->st));
		FG_CUDA(cudaStreamWaitEvent(f->st_up, f->ev_fence, 0));
		if (count) FG_CUDA(cudaMemcpyAsync(dst, host, count * 4, cudaMemcpyHostToDevice, f->st_up));
		if (field == FLIP_S) FG_TRY(queue_s_check(f, 0, count, true, f->st_up));
		FG_CUDA(cudaEventRecord(f->ev_uploaded, f->st_up));
		f->upload_pending = true;
		return FLIPGPU_OK;
	}
	FG_TRY(join_uploads(f));
	if (count) FG_CUDA(cudaMemcpyAsync(dst, host, count * 4, cudaMemcpyHostToDevice, f->st));
	// the blocked solvers pack s into one bit per neighbour; whether that is exact is checked on the device, behind the copy
	if (field == FLIP_S) FG_TRY(queue_s_check(f, 0, count, true, f->st));
	if (wait) FG_CUDA(cudaStreamSynchronize(f->st));
	return FLIPGPU_OK;
}
int flip_upload(FlipSim* f, int field, const void* host, size_t count) { return upload_impl(f, field, host, count, true); }
// The per-frame form (main.cpp:106-122 rewrites s, u, v every drag frame): the copy is only queued on the handle's stream.
// `host` must stay unchanged until the next synchronising call (flip_synchronize, flip_readback*, flip_readback_wait) unless it is
// pageable memory, which the CUDA runtime stages before returning.
int flip_upload_async(FlipSim* f, int field, const void* host, size_t count) { return upload_impl(f, field, host, count, false); }

int flip_readback(FlipSim* f, int field, void* host, size_t count) {
	FG_CHECK(f && host, FLIPGPU_EINVAL, "flip_readback: null argument");
	FG_CHECK(field >= 0 && field < FLIP_FIELD_COUNT, FLIPGPU_EINVAL, "flip_readback: unknown field %d", field);
	DeviceGuard guard(f->device);
	size_t want = field_count(f, field);
	FG_CHECK(count == want, FLIPGPU_EINVAL, "flip_readback: field %d holds %zu elements, asked for %zu", field, want, count);
	const void* srcp = grid_field_ptr(f, field);
	FG_CHECK(srcp || !f->sharded, FLIPGPU_ESTATE, "sharded simulation: use flip_readback_particles / flip_readback_rows");
	FG_TRY(join_uploads(f));
	int np = f->d.num_particles, c = f->cur, n = 1 - c;
	unsigned gr = pgrid(f);
	if (!srcp) {
		switch (field) {
			case FLIP_PARTICLE_POS:
				if (f->identity) srcp = f->pos[c];
				else { if (np) { k_scatter_to_caller<float2, 1><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->pos[c], f->pos[n]); f->lc.n++; } srcp = f->pos[n]; }
				break;
			case FLIP_PARTICLE_VEL:
				if (f->identity) srcp = f->vel[c];
				else { if (np) { k_scatter_to_caller<float2, 1><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->vel[c], f->vel[n]); f->lc.n++; } srcp = f->vel[n]; }
				break;
			case FLIP_PARTICLE_COLOR:
				FG_TRY(ensure_colors(f));
				if (f->identity) srcp = f->col[c];
				else { if (np) { k_scatter_to_caller<float, 3><<<gr, kThreads, 0, f->st>>>(np, f->orig[c], f->col[c], f->col[n]); f->lc.n++; } srcp = f->col[n]; }
				break;
			case FLIP_CELL_COLOR: FG_TRY(ensure_cell_colors(f)); srcp = f->cell_color; break;
			case FLIP_NUM_CELL_PARTICLES: srcp = f->count; break;
			case FLIP_FIRST_CELL_PARTICLE: srcp = f->first; break;
			case FLIP_CELL_PARTICLE_IDS:
				// before the first binning the reference array is all zero (Q4 convention)
				if (!f->bins_valid && f->identity) { FG_CUDA(cudaMemsetAsync(f->src, 0, (size_t)np * 4, f->st)); srcp = f->src; }
				else srcp = f->orig[c];
				break;
		}
	}
	FG_CUDA(cudaGetLastError());
	if (count) FG_CUDA(cudaMemcpyAsync(host, srcp, count * 4, cudaMemcpyDeviceToHost, f->st));
	FG_CUDA(cudaStreamSynchronize(f->st));
	return sor_workspace_check(f->ws);
}

// particle_pos in caller order -> host, without blocking the caller or the simulation stream (frame pipeline, see FlipSim).
// If the last flip_step already staged this state only the PCIe copy is queued here; otherwise the un-permute runs now.
int flip_readback_pos_async(FlipSim* f, void* host, size_t count) {
	FG_CHECK(f && host, FLIPGPU_EINVAL, "flip_readback_pos_async: null argument");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "sharded simulation: slots carry GLOBAL particle indices, use flip_readback_particles");
	FG_CHECK(count == 2 * (size_t)f->d.num_particles, FLIPGPU_EINVAL, "flip_readback_pos_async: particle_pos holds %zu elements, asked for %zu",
	         2 * (size_t)f->d.num_particles, count);
	DeviceGuard guard(f->device);
	f->frame_mode = true;
	if (f->d.num_particles == 0) return FLIPGPU_OK;
	if (!f->staged_valid) FG_TRY(stage_positions(f));
	const int b = f->frame_idx;
	FG_CUDA(cudaStreamWaitEvent(f->st_copy, f->ev_staged[b], 0));
	FG_CUDA(cudaMemcpyAsync(host, f->staging[b], count * 4, cudaMemcpyDeviceToHost, f->st_copy));
	FG_CUDA(cudaEventRecord(f->ev_copied[b], f->st_copy));
	f->copy_pending[b] = true;
	return FLIPGPU_OK;
}
int flip_readback_wait(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DeviceGuard guard(f->device);
	for (int b = 0; b < 2; b++) {
		if (f->copy_pending[b]) FG_CUDA(cudaEventSynchronize(f->ev_copied[b]));
		f->copy_pending[b] = false;
	}
	// The frame is complete on the host.  The step that produced it may still be running its solve (the positions are final
	// before it): a solver time-out of that step is reported by the next call that synchronises the simulation stream.
	return sor_workspace_check(f->ws);
}

int flip_set_obstacle(FlipSim* f, float x, float y, float vx, float vy, float radius) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DeviceGuard guard(f->device);
	Geo g = geo(f);
	FG_TRY(join_uploads(f));
	if (g.nx > 3 && g.ny > 3) {
		dim3 block(32, 8), grid(cdiv(g.nx - 3, 32), cdiv(g.ny - 3, 8));
		k_set_obstacle<<<grid, block, 0, f->st>>>(g, x, y, vx, vy, radius, f->s, f->u, f->v);
		f->lc.n++;
		FG_CUDA(cudaGetLastError());
	}
	return FLIPGPU_OK;
}

// one simulate(), flip_fluid.h:566-580 (Q11: one sub-step), queued on the handle's stream
static int one_step(FlipSim* f, const FlipStepParams* sp, float cp, bool stage_frame) {
	phase_mark(f, FLIP_PH_INTEGRATE_BIN);
	FG_TRY(integrate_sorted(f, sp->dt, sp->gravity));
	phase_mark(f, FLIP_PH_SEPARATE);
	// the last sweep runs the collision / marking pass itself (same device functions, identical results)
	const bool fused = sp->separate_particles && sp->num_particle_iters >= 1 && !f->exact_order && f->d.num_particles > 0;
	const FuseCollide fc{sp->obstacle_x, sp->obstacle_y, sp->obstacle_vel_x, sp->obstacle_vel_y, sp->obstacle_radius};
	if (sp->separate_particles) FG_TRY(separate(f, sp->num_particle_iters, sp->update_colors != 0, fused ? &fc : nullptr));
	phase_mark(f, FLIP_PH_COLLIDE_MARK);
	FG_TRY(collide_mark(f, true, true, sp->obstacle_x, sp->obstacle_y, sp->obstacle_vel_x, sp->obstacle_vel_y, sp->obstacle_radius, fused));
	phase_mark(f, FLIP_PH_P2G);
	FG_TRY(p2g(f));
	phase_mark(f, FLIP_PH_SOLVE);
	FG_TRY(solve(f, sp->num_pressure_iters, cp, sp->over_relaxation, sp->compensate_drift != 0, sp->solver_order));
	phase_mark(f, FLIP_PH_G2P);
	FG_TRY(g2p(f, sp->flip_ratio, stage_frame));   // frame mode: + the caller-order copy of the positions
	phase_mark(f, FLIP_PH_COLORS);
	if (sp->update_colors) FG_TRY(colors(f));
	phase_mark(f, FLIP_PH_COUNT);
	return FLIPGPU_OK;
}

// Replay (or first capture) of the whole-step graph for the current ping-pong parity.  Capture runs one_step() with the
// stream in capture mode: nothing executes, the host-side state advances exactly as in an eager step, and the graph is then
// launched once.  A replay launches the stored graph and applies the same host-side state changes.
static int graph_step(FlipSim* f, const FlipStepParams* sp, float cp) {
	FlipSim::StepGraph& G = f->graphs[f->cur];
	const bool fresh = !G.exec || memcmp(&G.sp, sp, sizeof(*sp)) != 0 || G.np != f->d.num_particles || G.s_binary != f->s_binary;
	if (fresh) {
		if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
		cudaGraph_t graph = nullptr;
		const long long l0 = f->lc.n;
		FG_CUDA(cudaStreamBeginCapture(f->st, cudaStreamCaptureModeThreadLocal));
		const int rc = one_step(f, sp, cp, false);
		const cudaError_t e = cudaStreamEndCapture(f->st, &graph);
		if (rc != FLIPGPU_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
		FG_CUDA(e);
		const cudaError_t ei = cudaGraphInstantiate(&G.exec, graph, 0);
		cudaGraphDestroy(graph);
		FG_CUDA(ei);
		G.sp = *sp; G.np = f->d.num_particles; G.s_binary = f->s_binary; G.launches = f->lc.n - l0;
		FG_CUDA(cudaGraphLaunch(G.exec, f->st));
		return FLIPGPU_OK;
	}
	FG_CUDA(cudaGraphLaunch(G.exec, f->st));
	f->cur = 1 - f->cur;            // what integrate_and_bin / positions_will_change do on the host
	f->identity = false; f->bins_valid = true; f->staged_valid = false;
	f->lc.n += G.launches;
	f->graph_replays++;
	return FLIPGPU_OK;
}

int flip_step(FlipSim* f, const FlipStepParams* sp, int n_steps) {
	FG_CHECK(f && sp, FLIPGPU_EINVAL, "flip_step: null argument");
	FG_CHECK(sp->num_pressure_iters >= 0 && sp->num_particle_iters >= 0 && sp->dt > 0, FLIPGPU_EINVAL, "flip_step: bad params");
	DeviceGuard guard(f->device);
	const float cp = f->d.density * f->d.h / sp->dt;  // flip_fluid.h:456
	if (f->sharded) {
		for (int k = 0; k < n_steps; k++) FG_TRY(shard_step(f, sp));
		return FLIPGPU_OK;
	}
	if (memcmp(&f->last_sp, sp, sizeof(*sp)) != 0) { f->last_sp = *sp; f->eager_steps = 0; }
	for (int k = 0; k < n_steps; k++) {
		if (f->timings_on) collect_slot(f, f->ev_slot);
		if (!f->rest_seen && f->h_pinned[0] != 0.f) f->rest_seen = true;
		const bool stage_frame = f->frame_mode && k == n_steps - 1;
		// steady state (allocations made, solver planned, rest density latched, nothing pending from the host): replay the graph
		const bool steady = f->graphs_on && f->eager_steps >= 2 && f->rest_seen && !f->timings_on && !f->exact_order && !f->s_check_pending &&
		                    !f->upload_pending && !stage_frame && sp->solver_order == FLIP_SOLVER_LEX_WAVEFRONT && f->s_binary && f->d.num_particles > 0;
		if (steady) {
			FG_TRY(graph_step(f, sp, cp));
			continue;
		}
		FG_TRY(one_step(f, sp, cp, stage_frame));
		f->eager_steps++;
		if (f->timings_on) {
			f->ev_pending[f->ev_slot] = true;
			f->ev_slot = (f->ev_slot + 1) % FlipSim::kEvRing;
		}
	}
	return FLIPGPU_OK;
}

int flip_set_graph_mode(FlipSim* f, int on) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	f->graphs_on = on != 0;
	return FLIPGPU_OK;
}
int flip_get_graph_replays(FlipSim* f, long long* replays) {
	FG_CHECK(f && replays, FLIPGPU_EINVAL, "null argument");
	*replays = f->graph_replays;
	return FLIPGPU_OK;
}


// ------------------------------------------------------------------ slab-sharded FLIP
int flip_create_sharded(const FlipParams* pr, int global_particles, int rank, int nranks, const void* unique_id128, FlipSim** out) {
	return flip_create_sharded_rows(pr, global_particles, rank, nranks, nullptr, unique_id128, out);
}
// row_starts: nranks+1 ascending row numbers (0 ... f_num_y), rank r owns rows [row_starts[r], row_starts[r+1]); NULL = equal slabs.
// Unequal slabs balance a scene whose liquid does not fill the partition axis evenly (BASELINE configs[3]: the top slabs are air).
int flip_create_sharded_rows(const FlipParams* pr, int global_particles, int rank, int nranks, const int* row_starts, const void* unique_id128,
                             FlipSim** out) {
	FG_CHECK(pr && out && nranks >= 1 && rank >= 0 && rank < nranks && global_particles >= 0, FLIPGPU_EINVAL, "flip_create_sharded: bad argument");
	FG_CHECK(nranks == 1 || unique_id128, FLIPGPU_EINVAL, "flip_create_sharded: a NCCL unique id is required for %d ranks", nranks);
	FlipSim* f = nullptr;
	FG_TRY(flip_create(pr, &f));
	DeviceGuard guard(f->device);
	const int ny = f->d.f_num_y;
	f->sharded = true; f->rank = rank; f->nranks = nranks;
	if (row_starts) {
		bool ok = row_starts[0] == 0 && row_starts[nranks] == ny;
		int min_rows = ny;
		for (int r = 0; r < nranks && ok; r++) { ok = row_starts[r + 1] > row_starts[r]; min_rows = std::min(min_rows, row_starts[r + 1] - row_starts[r]); }
		if (!ok) {
			flip_destroy(f);
			fg::set_error("flip_create_sharded_rows: row_starts must ascend from 0 to f_num_y = %d", ny);
			return FLIPGPU_EINVAL;
		}
		f->j0 = row_starts[rank]; f->j1 = row_starts[rank + 1]; f->min_slab_rows = min_rows;
	} else {
		f->j0 = (int)((long long)ny * rank / nranks);
		f->j1 = (int)((long long)ny * (rank + 1) / nranks);
		f->min_slab_rows = ny / nranks;
	}
	if (nranks > 1 && f->min_slab_rows < f->halo + 1) {
		flip_destroy(f);
		fg::set_error("flip_create_sharded: a slab of %d rows is thinner than the %d-row halo (%d rows over %d ranks)", f->min_slab_rows, f->halo, ny, nranks);
		return FLIPGPU_EINVAL;
	}
	f->gy0 = std::max(f->j0 - f->halo, 0); f->gy1 = std::min(f->j1 + f->halo, ny);
	f->hy0 = std::max((int)((float)f->gy0 * f->d.h * f->d.p_inv_spacing) - 1, 0);
	f->hy1 = std::min((int)((float)f->gy1 * f->d.h * f->d.p_inv_spacing) + 2, f->d.p_num_y);
	f->cap_send = std::max(f->d.max_particles / 2, 1024);
	f->chunk_stride = (3 * (size_t)cdiv((size_t)std::max(f->d.max_particles, 1), (size_t)kShardChunk) + 7) & ~(size_t)3;
	FG_CUDA(cudaMalloc(&f->d_chunk, 2 * f->chunk_stride * sizeof(int)));
	f->global_particles = global_particles;
	for (int k = 0; k < 2; k++) {
		FG_CUDA(cudaMalloc(&f->send_pos[k], (size_t)f->cap_send * sizeof(float2)));
		FG_CUDA(cudaMalloc(&f->send_vel[k], (size_t)f->cap_send * sizeof(float2)));
		FG_CUDA(cudaMalloc(&f->send_id[k], (size_t)f->cap_send * 4));
	}
	FG_CUDA(cudaMalloc(&f->d_counters, 12 * sizeof(int)));
	FG_CUDA(cudaMemset(f->d_counters, 0, 12 * sizeof(int)));
	FG_CUDA(cudaMallocHost(&f->h_counters, 8 * sizeof(int)));
	f->d.num_particles = 0;
	f->identity = false;
	if (nranks > 1) {
		FG_TRY(load_nccl());
		ncclUniqueId id;
		memcpy(&id, unique_id128, sizeof(id));
		FG_NCCL(g_nccl.CommInitRank(&f->comm, nranks, id, rank));
	}
	*out = f;
	return FLIPGPU_OK;
}

int flip_get_shard_info(FlipSim* f, int* j0, int* j1, int* halo_rows, int* ghost_band_rows, long long* exchanges) {
	FG_CHECK(f && j0 && j1 && halo_rows && ghost_band_rows && exchanges, FLIPGPU_EINVAL, "null argument");
	FG_CHECK(f->sharded, FLIPGPU_ESTATE, "not a sharded simulation");
	*j0 = f->j0; *j1 = f->j1; *halo_rows = f->halo; *ghost_band_rows = f->band; *exchanges = f->exchanges;
	return FLIPGPU_OK;
}

// this rank's particles with their GLOBAL caller indices (any order); replaces whatever the rank held
int flip_upload_particles(FlipSim* f, const float* pos, const float* vel, const int* ids, int n) {
	FG_CHECK(f && (n == 0 || (pos && ids)), FLIPGPU_EINVAL, "flip_upload_particles: null argument");
	FG_CHECK(f->sharded, FLIPGPU_ESTATE, "flip_upload_particles is for sharded simulations");
	FG_CHECK(n >= 0 && n <= f->d.max_particles, FLIPGPU_EINVAL, "flip_upload_particles: %d particles exceed the local capacity %d", n, f->d.max_particles);
	for (int k = 0; k < n; k++)
		FG_CHECK(ids[k] >= 0 && ids[k] < f->global_particles, FLIPGPU_EINVAL, "flip_upload_particles: id %d outside [0,%d)", ids[k], f->global_particles);
	DeviceGuard guard(f->device);
	const int c = f->cur;
	if (n) {
		FG_CUDA(cudaMemcpyAsync(f->pos[c], pos, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, f->st));
		if (vel) FG_CUDA(cudaMemcpyAsync(f->vel[c], vel, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, f->st));
		else FG_CUDA(cudaMemsetAsync(f->vel[c], 0, (size_t)n * sizeof(float2), f->st));
		FG_CUDA(cudaMemcpyAsync(f->orig[c], ids, (size_t)n * 4, cudaMemcpyHostToDevice, f->st));
	}
	FG_CUDA(cudaStreamSynchronize(f->st));
	f->d.num_particles = n; f->n_owned = n; f->bins_valid = false; f->identity = false; f->vmax_valid = false;
	return FLIPGPU_OK;
}

// the particles this rank OWNS now (by position), with their global indices; *n_out = how many
int flip_readback_particles(FlipSim* f, float* pos, float* vel, int* ids, int capacity, int* n_out) {
	FG_CHECK(f && n_out, FLIPGPU_EINVAL, "flip_readback_particles: null argument");
	FG_CHECK(f->sharded, FLIPGPU_ESTATE, "flip_readback_particles is for sharded simulations");
	DeviceGuard guard(f->device);
	Geo g = geo(f);
	const int c = f->cur, n = 1 - c;
	FG_CUDA(cudaMemsetAsync(f->d_counters, 0, 8 * sizeof(int), f->st));
	FG_TRY(shard_compact(f, false));
	FG_CUDA(cudaMemcpyAsync(f->h_counters, f->d_counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, f->st));
	FG_CUDA(cudaStreamSynchronize(f->st));
	const int k = f->h_counters[0];
	*n_out = k;
	FG_CHECK(k <= capacity, FLIPGPU_EINVAL, "flip_readback_particles: %d owned particles, capacity %d", k, capacity);
	if (k) {
		if (pos) FG_CUDA(cudaMemcpyAsync(pos, f->pos[n], (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, f->st));
		if (vel) FG_CUDA(cudaMemcpyAsync(vel, f->vel[n], (size_t)k * sizeof(float2), cudaMemcpyDeviceToHost, f->st));
		if (ids) FG_CUDA(cudaMemcpyAsync(ids, f->orig[n], (size_t)k * 4, cudaMemcpyDeviceToHost, f->st));
	}
	FG_CUDA(cudaStreamSynchronize(f->st));
	return FLIPGPU_OK;
}

// rows [j_start, j_start+n_rows) of a grid field (global row numbering, nx elements per row)
static int rows_copy(FlipSim* f, int field, void* host, int j_start, int n_rows, bool to_device) {
	FG_CHECK(f && host, FLIPGPU_EINVAL, "null argument");
	char* base = (char*)grid_field_ptr(f, field);
	FG_CHECK(base, FLIPGPU_EINVAL, "rows access is for grid fields (got %d)", field);
	FG_CHECK(j_start >= 0 && n_rows >= 0 && j_start + n_rows <= f->d.f_num_y, FLIPGPU_EINVAL, "rows [%d,%d) outside the grid", j_start, j_start + n_rows);
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	size_t off = (size_t)j_start * f->d.f_num_x * 4, bytes = (size_t)n_rows * f->d.f_num_x * 4;
	if (bytes) {
		if (to_device) FG_CUDA(cudaMemcpyAsync(base + off, host, bytes, cudaMemcpyHostToDevice, f->st));
		else FG_CUDA(cudaMemcpyAsync(host, base + off, bytes, cudaMemcpyDeviceToHost, f->st));
		if (to_device && field == FLIP_S) FG_TRY(queue_s_check(f, (size_t)j_start * f->d.f_num_x, (size_t)n_rows * f->d.f_num_x, false, f->st));
	}
	FG_CUDA(cudaStreamSynchronize(f->st));
	if (!to_device) FG_TRY(sor_workspace_check(f->ws));
	return FLIPGPU_OK;
}
int flip_upload_rows(FlipSim* f, int field, const void* host, int j_start, int n_rows) { return rows_copy(f, field, const_cast<void*>(host), j_start, n_rows, true); }
int flip_readback_rows(FlipSim* f, int field, void* host, int j_start, int n_rows) { return rows_copy(f, field, host, j_start, n_rows, false); }

int flip_phase_integrate(FlipSim* f, float dt, float gravity) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	return integrate_and_bin(f, true, false, dt, gravity);
}
int flip_phase_bin(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	return integrate_and_bin(f, false, true, 0.f, 0.f);
}
int flip_phase_separate(FlipSim* f, int num_iters, int update_colors) {
	FG_CHECK(f && num_iters >= 0, FLIPGPU_EINVAL, "flip_phase_separate: bad argument");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	if (!f->bins_valid) FG_TRY(integrate_and_bin(f, false, true, 0.f, 0.f));
	return separate(f, num_iters, update_colors != 0);
}
int flip_phase_collide(FlipSim* f, float ox, float oy, float ovx, float ovy, float orad) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	return collide_mark(f, true, false, ox, oy, ovx, ovy, orad);
}
int flip_phase_p2g(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	if (!f->rest_seen && f->h_pinned[0] != 0.f) f->rest_seen = true;
	FG_TRY(collide_mark(f, false, true, 0, 0, 0, 0, 0));
	return p2g(f);
}
int flip_phase_solve(FlipSim* f, int iters, float dt, float omega, int drift, int order) {
	FG_CHECK(f && dt > 0 && iters >= 0, FLIPGPU_EINVAL, "flip_phase_solve: bad argument");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	// :452-454
	FG_CUDA(cudaMemsetAsync(f->p, 0, (size_t)f->d.f_num_cells * 4, f->st));
	FG_CUDA(cudaMemcpyAsync(f->prev_u, f->u, (size_t)f->d.f_num_cells * 4, cudaMemcpyDeviceToDevice, f->st));
	FG_CUDA(cudaMemcpyAsync(f->prev_v, f->v, (size_t)f->d.f_num_cells * 4, cudaMemcpyDeviceToDevice, f->st));
	return solve(f, iters, f->d.density * f->d.h / dt, omega, drift != 0, order);
}
int flip_phase_g2p(FlipSim* f, float flip_ratio) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	return g2p(f, flip_ratio);
}
int flip_phase_colors(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded, FLIPGPU_ESTATE, "flip_phase_* are single-GPU replay entry points (a sharded step needs its exchanges: flip_step)");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	return colors(f);
}

int flip_get_scalar(FlipSim* f, int which, float* out) {
	FG_CHECK(f && out, FLIPGPU_EINVAL, "flip_get_scalar: null argument");
	DeviceGuard guard(f->device);
	switch (which) {
		case FLIP_SCALAR_REST_DENSITY:
			FG_CUDA(cudaMemcpyAsync(out, f->d_rest, 4, cudaMemcpyDeviceToHost, f->st));
			FG_CUDA(cudaStreamSynchronize(f->st));
			if (*out != 0.f) f->rest_seen = true;
			return FLIPGPU_OK;
		case FLIP_SCALAR_RESIDUAL_MAX:
		case FLIP_SCALAR_RESIDUAL_RMS: {
			float r2[2];
			FG_TRY(join_uploads(f));
			FG_TRY(sor_residual(sor_view(f), true, f->d_partials, kPartials, r2, f->st, &f->lc));
			*out = which == FLIP_SCALAR_RESIDUAL_MAX ? r2[0] : r2[1];
			return FLIPGPU_OK;
		}
	}
	fg::set_error("flip_get_scalar: unknown scalar %d", which);
	return FLIPGPU_EINVAL;
}

int flip_set_rest_density(FlipSim* f, float rest) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DeviceGuard guard(f->device);
	FG_CUDA(cudaMemcpyAsync(f->d_rest, &rest, 4, cudaMemcpyHostToDevice, f->st));
	FG_CUDA(cudaStreamSynchronize(f->st));
	f->rest_seen = rest != 0.f;
	f->h_pinned[0] = rest;
	return FLIPGPU_OK;
}

int flip_set_exact_order(FlipSim* f, int on) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	FG_CHECK(!f->sharded || !on, FLIPGPU_ESTATE, "the reference-order mode is a single-GPU parity mode");
	f->exact_order = on != 0;
	return FLIPGPU_OK;
}
int flip_get_exact_order_info(FlipSim* f, int* levels, int* drifted) {
	FG_CHECK(f && levels && drifted, FLIPGPU_EINVAL, "null argument");
	DeviceGuard guard(f->device);
	*levels = f->sep_levels; *drifted = 0;
	if (f->d_flags) {
		FG_CUDA(cudaMemcpyAsync(drifted, f->d_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, f->st));
		FG_CUDA(cudaStreamSynchronize(f->st));
	}
	return FLIPGPU_OK;
}
int flip_synchronize(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	DeviceGuard guard(f->device);
	FG_TRY(join_uploads(f));
	FG_CUDA(cudaStreamSynchronize(f->st));
	return sor_workspace_check(f->ws);  // FLIPGPU_ESTATE if a solver dependency wait timed out since the last check
}
int flip_get_stream(FlipSim* f, void** stream) {
	FG_CHECK(f && stream, FLIPGPU_EINVAL, "null argument");
	*stream = (void*)f->st;
	return FLIPGPU_OK;
}
int flip_enable_timings(FlipSim* f, int on) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	if (!on) collect_timings(f);
	f->timings_on = on != 0;
	return FLIPGPU_OK;
}
int flip_reset_timings(FlipSim* f) {
	FG_CHECK(f, FLIPGPU_EINVAL, "null handle");
	collect_timings(f);
	for (float& m : f->ph_ms) m = 0.f;
	return FLIPGPU_OK;
}
int flip_get_timings(FlipSim* f, float* ms, int count) {
	FG_CHECK(f && ms && count >= 0, FLIPGPU_EINVAL, "flip_get_timings: bad argument");
	collect_timings(f);
	for (int i = 0; i < count && i < FLIP_PH_COUNT; i++) ms[i] = f->ph_ms[i];
	return FLIPGPU_OK;
}
int flip_debug_solver_profile(FlipSim* f, unsigned long long* out8) {
	FG_CHECK(f && out8, FLIPGPU_EINVAL, "null argument");
	DeviceGuard guard(f->device);
	FG_CUDA(cudaStreamSynchronize(f->st));
	return sor_workspace_profile(f->ws, out8);
}
int flip_get_launch_count(FlipSim* f, long long* launches) {
	FG_CHECK(f && launches, FLIPGPU_EINVAL, "null argument");
	*launches = f->lc.n;
	return FLIPGPU_OK;
}

// flip_fluid/src/main.cpp:47-77 -- particle lattice and tank walls, host fp32 arithmetic as written there
int flip_scene_tank(float tank_width, float tank_height, float spacing, float radius, float rel_w, float rel_h,
                    int f_num_x, int f_num_y, float* pos, size_t pos_capacity, float* s, int* num_particles) {
	FG_CHECK(num_particles, FLIPGPU_EINVAL, "flip_scene_tank: num_particles is null");
	float h = spacing, r = radius;
	float dx = 2 * r;
	float dy = std::sqrt(3.0) / 2 * dx;  // :49 goes through double
	int num_x = (rel_w * tank_width - 2 * h - 2 * r) / dx;
	int num_y = (rel_h * tank_height - 2 * h - 2 * r) / dy;
	if (num_x < 0) num_x = 0;
	if (num_y < 0) num_y = 0;
	long long total = (long long)num_x * num_y;
	FG_CHECK(total < (1ll << 31), FLIPGPU_EINVAL, "flip_scene_tank: %lld particles overflow int32", total);
	*num_particles = (int)total;
	if (pos) {
		FG_CHECK(pos_capacity >= 2 * (size_t)total, FLIPGPU_EINVAL, "flip_scene_tank: pos buffer too small");
		size_t p = 0;
		for (int i = 0; i < num_x; i++)
			for (int j = 0; j < num_y; j++) {
				pos[p++] = h + r + dx * i + (j % 2 == 0 ? 0 : r);
				pos[p++] = h + r + dy * j;
			}
	}
	if (s)
		for (int j = 0; j < f_num_y; j++)
			for (int i = 0; i < f_num_x; i++)
				s[i + (size_t)f_num_x * j] = (i == 0 || i == f_num_x - 1 || j == 0 || j == f_num_y - 1) ? 0.f : 1.f;
	return FLIPGPU_OK;
}

}  // extern "C"
