// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Headless harness around the UNMODIFIED reference header
//   /root/reference/eulerian_fluid/src/fluid3d.h   (struct Fluid3D, lines 19-368)
// compiled where it lies (own translation unit: the header defines global min/max templates, fluid3d.h:8-16).
// Built by oracle/Makefile into oracle/_ref/libref_fluid3d.so (g++ -O2 -std=c++17 -include cstring).
// new_u/new_v/new_w/pressure/new_m are zero-filled after construction (the reference leaves them uninitialised, :46-60).
// fluid3d.h:311 bounds advectVel's k loop by num_y: with num_z < num_y the reference reads out of bounds, so the tests
// only drive this harness with num_z >= num_y.
#include <cstring>
#include "fluid3d.h"

extern "C" {

void* reffluid3d_create(int x, int y, int z, float density, float h) {
	Fluid3D* f = new Fluid3D(x, y, z, density, h);
	size_t C = f->num_cells;
	std::memset(f->new_u, 0, 4 * C); std::memset(f->new_v, 0, 4 * C); std::memset(f->new_w, 0, 4 * C);
	std::memset(f->pressure, 0, 4 * C); std::memset(f->new_m, 0, 4 * C);
	return f;
}
void reffluid3d_destroy(void* h) { delete (Fluid3D*)h; }
void reffluid3d_dims(void* h, int* ints, float* floats) {
	Fluid3D* f = (Fluid3D*)h;
	ints[0] = f->num_x; ints[1] = f->num_y; ints[2] = f->num_z; ints[3] = f->num_cells;
	floats[0] = f->density; floats[1] = f->h;
}
void* reffluid3d_ptr(void* h, const char* name) {
	Fluid3D* f = (Fluid3D*)h;
#define FIELD(n) if (!std::strcmp(name, #n)) return (void*)f->n;
	FIELD(u) FIELD(v) FIELD(w) FIELD(new_u) FIELD(new_v) FIELD(new_w) FIELD(pressure) FIELD(solid) FIELD(m) FIELD(new_m)
#undef FIELD
	return nullptr;
}
void reffluid3d_solve(void* h, int iters, float dt) { ((Fluid3D*)h)->solveIncompressibility(iters, dt); }
void reffluid3d_extrapolate(void* h) { ((Fluid3D*)h)->extrapolate(); }
void reffluid3d_advect_vel(void* h, float dt) { ((Fluid3D*)h)->advectVel(dt); }
void reffluid3d_advect_smoke(void* h, float dt) { ((Fluid3D*)h)->advectSmoke(dt); }
float reffluid3d_sample(void* h, float x, float y, float z, int field) { return ((Fluid3D*)h)->sampleField(x, y, z, field); }
void reffluid3d_step(void* h, int iters, float dt, int n_steps) {
	Fluid3D* f = (Fluid3D*)h;
	for (int k = 0; k < n_steps; k++) {
		f->solveIncompressibility(iters, dt);
		f->extrapolate();
		f->advectVel(dt);
		f->advectSmoke(dt);
	}
}
}
