// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Headless harness around the UNMODIFIED reference header
//   /root/reference/eulerian_fluid/src/fluid.h   (class Fluid, lines 19-260)
// compiled where it lies.  Own translation unit because fluid.h defines global
// min/max templates (fluid.h:9-17).  Built by oracle/Makefile into
// oracle/_ref/libref_euler.so (g++ -O2 -std=c++17 -include cstring).
// new_u/new_v/pressure/new_m are zero-filled after construction (the reference
// leaves them uninitialised, fluid.h:50-68).
#include <cstring>
#include <chrono>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif
#include "fluid.h"

extern "C" {

void* refeuler_create(int x, int y, float density, float h) {
	Fluid* f = new Fluid(x, y, density, h);
	size_t C = f->getNumCells();
	std::memset(f->new_u, 0, 4 * C); std::memset(f->new_v, 0, 4 * C);
	std::memset(f->pressure, 0, 4 * C); std::memset(f->new_m, 0, 4 * C);
	return f;
}
void refeuler_destroy(void* h) { delete (Fluid*)h; }
void refeuler_dims(void* h, int* ints, float* floats) {
	Fluid* f = (Fluid*)h;
	ints[0] = f->getNumX(); ints[1] = f->getNumY(); ints[2] = f->getNumCells();
	floats[0] = f->density; floats[1] = f->h; floats[2] = f->h1;
}
void* refeuler_ptr(void* h, const char* name) {
	Fluid* f = (Fluid*)h;
#define FIELD(n) if (!std::strcmp(name, #n)) return (void*)f->n;
	FIELD(u) FIELD(v) FIELD(new_u) FIELD(new_v) FIELD(pressure) FIELD(solid) FIELD(m) FIELD(new_m)
#undef FIELD
	return nullptr;
}
void refeuler_solve(void* h, int iters, float dt) { ((Fluid*)h)->solveIncompressibility(iters, dt); }
void refeuler_extrapolate(void* h) { ((Fluid*)h)->extrapolate(); }
void refeuler_advect_vel(void* h, float dt) { ((Fluid*)h)->advectVel(dt); }
void refeuler_advect_smoke(void* h, float dt) { ((Fluid*)h)->advectSmoke(dt); }
float refeuler_sample(void* h, float x, float y, int field) { return ((Fluid*)h)->sampleField(x, y, field); }
// FluidUI::setObstacle (eulerian_fluid/src/fluid_ui.h:24-62) lives in the UI shell, whose header needs the windowing
// engine; this is the harness's restatement of it acting on the reference object's public arrays, so that scenes with
// a dragged obstacle can be driven through the reference solver.  (vx, vy) = (x - previous x)/dt or 0 (:39-43).
void refeuler_set_obstacle(void* h, int x, int y, int r, float vx, float vy) {
	Fluid* f = (Fluid*)h;
	const int nx = f->getNumX(), ny = f->getNumY();
	std::memset(f->solid, false, sizeof(bool) * nx * ny);
	for (int j = 0; j < ny; j++) f->solid[f->ix(0, j)] = true;
	for (int i = 0; i < nx; i++) { f->solid[f->ix(i, 0)] = true; f->solid[f->ix(i, ny - 1)] = true; }
	for (int i = 1; i < nx - 1; i++)
		for (int j = 1; j < ny - 1; j++) {
			int dx = i - x, dy = j - y;
			if (dx * dx + dy * dy < r * r) {
				int c = f->ix(i, j);
				f->solid[c] = true; f->m[c] = 1.f; f->u[c] = vx; f->v[c] = vy;
			}
		}
}
// on x86 the reference's solve stalls on denormals over large quiescent regions.  Flip the calling
// thread's MXCSR flush-to-zero (bit 15) + denormals-are-zero (bit 6) so the bench can time the reference both ways;
// returns the previous setting (0/1), -1 where MXCSR does not exist.  Timing aid only: parity runs never enable it.
int refeuler_set_ftz_daz(int on) {
#if defined(__x86_64__) || defined(__i386__)
	unsigned csr = _mm_getcsr();
	int was = (csr & 0x8040u) == 0x8040u;
	_mm_setcsr(on ? (csr | 0x8040u) : (csr & ~0x8040u));
	return was;
#else
	(void)on;
	return -1;
#endif
}
// caller sequence of eulerian_fluid/src/fluid_ui.h:107-111
void refeuler_step(void* h, int iters, float dt, int n_steps, double* phase_ns) {
	Fluid* f = (Fluid*)h;
	using clk = std::chrono::steady_clock;
	for (int k = 0; k < n_steps; k++) {
		auto t0 = clk::now();
		f->solveIncompressibility(iters, dt);
		auto t1 = clk::now();
		f->extrapolate();
		f->advectVel(dt);
		f->advectSmoke(dt);
		auto t2 = clk::now();
		if (phase_ns) {
			phase_ns[0] += std::chrono::duration<double, std::nano>(t1 - t0).count();
			phase_ns[1] += std::chrono::duration<double, std::nano>(t2 - t1).count();
		}
	}
}
}
