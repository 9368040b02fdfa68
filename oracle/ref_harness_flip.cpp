// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Headless harness around the UNMODIFIED reference header
//   /root/reference/flip_fluid/src/flip_fluid.h   (struct FlipFluid, lines 13-582)
// compiled where it lies (no sources are copied into this repo).  Built by
// oracle/Makefile into oracle/_ref/libref_flip.so with
//   g++ -O2 -std=c++17 -include cstring   (the header forgets <cstring>)
// and no -march=native / -ffast-math, so fp32 stays unfused like MSVC /fp:precise.
//
// The harness only adds what the reference leaves undefined:
//   * Q4: every state array is zero-filled after construction (the reference
//     allocates with new[] and never initialises; flip_fluid.h:76-110).
//   * Q3: cell_type is re-pointed into a buffer with f_num_x+1 guard ints
//     (= SolidCell) in front, because flip_fluid.h:406-409 reads
//     cell_type[nr-offset] below the array when y0==0.
// Every phase of simulate() (flip_fluid.h:560-581) is exported separately so
// tests can replay one kernel at a time from identical state.
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <chrono>
#include "flip_fluid.h"

namespace {
struct Harness {
	FlipFluid* f = nullptr;
	int* orig_cell_type = nullptr;  // what the reference allocated (restored before delete)
	int* guarded = nullptr;
	double phase_ns[10] = {0};
};
template <class T> void zero(T* p, size_t n) { std::memset(p, 0, sizeof(T) * n); }
using clk = std::chrono::steady_clock;
inline double ns_since(clk::time_point t0) {
	return std::chrono::duration<double, std::nano>(clk::now() - t0).count();
}
}  // namespace

extern "C" {

void* refflip_create(float density, float width, float height, float spacing, float radius, int max_particles) {
	Harness* hs = new Harness;
	hs->f = new FlipFluid(density, width, height, spacing, radius, max_particles);
	FlipFluid& f = *hs->f;
	size_t C = f.f_num_cells;
	zero(f.u, C); zero(f.v, C); zero(f.du, C); zero(f.dv, C);
	zero(f.prev_u, C); zero(f.prev_v, C); zero(f.p, C); zero(f.s, C);
	zero(f.cell_color, 3 * C); zero(f.particle_density, C);
	zero(f.particle_pos, 2 * (size_t)max_particles);
	zero(f.particle_vel, 2 * (size_t)max_particles);
	for (int i = 0; i < max_particles; i++) { f.particle_color[3 * i] = 0; f.particle_color[3 * i + 1] = 0; }
	zero(f.num_cell_particles, f.p_num_cells);
	zero(f.first_cell_particle, f.p_num_cells + 1);
	zero(f.cell_particle_ids, max_particles);
	// Q3 guard
	int guard = f.f_num_x + 1;
	hs->guarded = (int*)std::malloc(sizeof(int) * (C + guard));
	for (int i = 0; i < guard; i++) hs->guarded[i] = FlipFluid::SolidCell;
	zero(hs->guarded + guard, C);
	hs->orig_cell_type = f.cell_type;
	f.cell_type = hs->guarded + guard;
	return hs;
}

void refflip_destroy(void* h) {
	Harness* hs = (Harness*)h;
	hs->f->cell_type = hs->orig_cell_type;
	delete hs->f;
	std::free(hs->guarded);
	delete hs;
}

// ints: f_num_x f_num_y f_num_cells p_num_x p_num_y p_num_cells max_particles num_particles
// floats: density h f_inv_spacing particle_radius p_inv_spacing particle_rest_density
void refflip_dims(void* h, int* ints, float* floats) {
	FlipFluid& f = *((Harness*)h)->f;
	ints[0] = f.f_num_x; ints[1] = f.f_num_y; ints[2] = f.f_num_cells;
	ints[3] = f.p_num_x; ints[4] = f.p_num_y; ints[5] = f.p_num_cells;
	ints[6] = f.max_particles; ints[7] = f.num_particles;
	floats[0] = f.density; floats[1] = f.h; floats[2] = f.f_inv_spacing;
	floats[3] = f.particle_radius; floats[4] = f.p_inv_spacing; floats[5] = f.particle_rest_density;
}

void* refflip_ptr(void* h, const char* name) {
	FlipFluid& f = *((Harness*)h)->f;
#define FIELD(n) if (!std::strcmp(name, #n)) return (void*)f.n;
	FIELD(u) FIELD(v) FIELD(du) FIELD(dv) FIELD(prev_u) FIELD(prev_v) FIELD(p) FIELD(s)
	FIELD(cell_type) FIELD(cell_color) FIELD(particle_pos) FIELD(particle_color) FIELD(particle_vel)
	FIELD(particle_density) FIELD(num_cell_particles) FIELD(first_cell_particle) FIELD(cell_particle_ids)
#undef FIELD
	return nullptr;
}

void refflip_set_num_particles(void* h, int n) { ((Harness*)h)->f->num_particles = n; }
void refflip_set_rest_density(void* h, float d) { ((Harness*)h)->f->particle_rest_density = d; }

void refflip_integrate(void* h, float dt, float g) { ((Harness*)h)->f->integrateParticles(dt, g); }
void refflip_push_apart(void* h, int iters) { ((Harness*)h)->f->pushParticlesApart(iters); }
void refflip_collide(void* h, float ox, float oy, float ovx, float ovy, float orad) {
	((Harness*)h)->f->handleParticleCollisions(ox, oy, ovx, ovy, orad);
}
void refflip_transfer(void* h, int to_grid, float flip_ratio) { ((Harness*)h)->f->transferVelocities(to_grid != 0, flip_ratio); }
void refflip_density(void* h) { ((Harness*)h)->f->updateParticleDensity(); }
void refflip_solve(void* h, int iters, float dt, float omega, int drift) {
	((Harness*)h)->f->solveIncompressibility(iters, dt, omega, drift != 0);
}
void refflip_particle_colors(void* h) { ((Harness*)h)->f->updateParticleColors(); }
void refflip_cell_colors(void* h) { ((Harness*)h)->f->updateCellColors(); }

void refflip_simulate(void* h, float dt, float gravity, float flip_ratio, int num_pressure_iters,
                      int num_particle_iters, float over_relaxation, int compensate_drift, int separate_particles,
                      float ox, float oy, float ovx, float ovy, float orad, int n_steps) {
	FlipFluid& f = *((Harness*)h)->f;
	for (int k = 0; k < n_steps; k++)
		f.simulate(dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
		           compensate_drift != 0, separate_particles != 0, ox, oy, ovx, ovy, orad);
}

// Same sequence as simulate() (flip_fluid.h:566-580) with steady_clock around each phase.
// phase_ns: 0 integrate 1 pushApart 2 collide 3 P2G 4 density 5 solve 6 G2P 7 colours
void refflip_simulate_timed(void* h, float dt, float gravity, float flip_ratio, int num_pressure_iters,
                            int num_particle_iters, float over_relaxation, int compensate_drift, int separate_particles,
                            float ox, float oy, float ovx, float ovy, float orad, int n_steps, double* phase_ns) {
	FlipFluid& f = *((Harness*)h)->f;
	for (int k = 0; k < n_steps; k++) {
		auto t = clk::now(); f.integrateParticles(dt, gravity); phase_ns[0] += ns_since(t);
		t = clk::now(); if (separate_particles) f.pushParticlesApart(num_particle_iters); phase_ns[1] += ns_since(t);
		t = clk::now(); f.handleParticleCollisions(ox, oy, ovx, ovy, orad); phase_ns[2] += ns_since(t);
		t = clk::now(); f.transferVelocities(true); phase_ns[3] += ns_since(t);
		t = clk::now(); f.updateParticleDensity(); phase_ns[4] += ns_since(t);
		t = clk::now(); f.solveIncompressibility(num_pressure_iters, dt, over_relaxation, compensate_drift != 0); phase_ns[5] += ns_since(t);
		t = clk::now(); f.transferVelocities(false, flip_ratio); phase_ns[6] += ns_since(t);
		t = clk::now(); f.updateParticleColors(); f.updateCellColors(); phase_ns[7] += ns_since(t);
	}
}

}  // extern "C"
