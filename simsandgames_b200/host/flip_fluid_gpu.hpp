// C++ facades over the C ABI (include/flipgpu.h) with the reference's member names, so the headless
// driver (headless.cpp) is `template <class Sim> run(...)` over either implementation:
//   FlipFluidGPU  <->  struct FlipFluid   reference flip_fluid/src/flip_fluid.h:13-582
//   FluidGPU      <->  class  Fluid       reference eulerian_fluid/src/fluid.h:19-260
// The reference exposes raw public arrays that its UI shell reads and writes between steps
// (flip_fluid/src/main.cpp:59-77,106-122,173-194).  Here the arrays live in HBM; the facade keeps host
// mirrors of exactly those caller-touched arrays:
//   written by the caller : particle_pos, num_particles, s, u, v        -> push() (called by simulate() when dirty)
//   read by the caller    : particle_pos, particle_color, cell_color    -> pull_render_state()
// Everything else is reachable through download(FLIP_xxx, ...).  Errors throw std::runtime_error with
// flipgpu_last_error(); there is no CPU fallback.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/flipgpu.h"

namespace flipgpu {

inline void check(int rc, const char* what) {
	if (rc != FLIPGPU_OK) throw std::runtime_error(std::string(what) + ": " + flipgpu_last_error());
}

struct FlipFluidGPU {
	// ---- the reference's public scalars (flip_fluid.h:14-57)
	float density = 0;
	int f_num_x = 0, f_num_y = 0, f_num_cells = 0;
	float h = 0, f_inv_spacing = 0;
	int max_particles = 0;
	float particle_radius = 0, p_inv_spacing = 0;
	int p_num_x = 0, p_num_y = 0, p_num_cells = 0;
	int num_particles = 0;  // the caller assigns this directly (main.cpp:59)
	enum CellType { FluidCell = 0, AirCell, SolidCell };
	// ---- host mirrors of the caller-touched arrays (reference layouts)
	std::vector<float> particle_pos, particle_color, s, u, v, cell_color;
	// extra switches of this implementation
	int solver_order = FLIP_SOLVER_LEX_WAVEFRONT;
	bool update_colors = true;

	FlipFluidGPU(float d, float width, float height, float sp, float r, int m, int device = -1) {
		FlipParams p{d, width, height, sp, r, m, device};
		check(flip_create(&p, &sim_), "flip_create");
		FlipDims dm;
		check(flip_get_dims(sim_, &dm), "flip_get_dims");
		density = dm.density; f_num_x = dm.f_num_x; f_num_y = dm.f_num_y; f_num_cells = dm.f_num_cells;
		h = dm.h; f_inv_spacing = dm.f_inv_spacing; max_particles = dm.max_particles;
		particle_radius = dm.particle_radius; p_inv_spacing = dm.p_inv_spacing;
		p_num_x = dm.p_num_x; p_num_y = dm.p_num_y; p_num_cells = dm.p_num_cells;
		particle_pos.assign(2 * (size_t)m, 0.f);
		particle_color.assign(3 * (size_t)m, 0.f);
		for (int i = 0; i < m; i++) particle_color[2 + 3 * (size_t)i] = 1;  // flip_fluid.h:94-96
		s.assign(f_num_cells, 0.f); u.assign(f_num_cells, 0.f); v.assign(f_num_cells, 0.f);
		cell_color.assign(3 * (size_t)f_num_cells, 0.f);
	}
	FlipFluidGPU(const FlipFluidGPU&) = delete;             // flip_fluid.h:116
	FlipFluidGPU& operator=(const FlipFluidGPU&) = delete;  // flip_fluid.h:145
	~FlipFluidGPU() { flip_destroy(sim_); }

	int fIX(int i, int j) const { return i + f_num_x * j; }  // flip_fluid.h:147-149
	int pIX(int i, int j) const { return i + p_num_x * j; }  // flip_fluid.h:151-153

	// the caller changed particle_pos / num_particles / s / u / v on the host
	void mark_dirty(bool particles = true, bool grid = true) { dirty_particles_ |= particles; dirty_grid_ |= grid; }

	void push() {
		if (dirty_particles_) {
			check(flip_set_num_particles(sim_, num_particles), "flip_set_num_particles");
			check(flip_upload(sim_, FLIP_PARTICLE_POS, particle_pos.data(), 2 * (size_t)num_particles), "upload particle_pos");
		}
		if (dirty_grid_) {
			check(flip_upload(sim_, FLIP_S, s.data(), s.size()), "upload s");
			check(flip_upload(sim_, FLIP_U, u.data(), u.size()), "upload u");
			check(flip_upload(sim_, FLIP_V, v.data(), v.size()), "upload v");
		}
		dirty_particles_ = dirty_grid_ = false;
	}

	// FlipFluid::simulate, flip_fluid.h:560-565 -- same 13 arguments
	void simulate(float dt, float gravity, float flip_ratio, int num_pressure_iters, int num_particle_iters,
	              float over_relaxation, bool compensate_drift, bool separate_particles, float obstacle_x, float obstacle_y,
	              float obstacle_vel_x, float obstacle_vel_y, float obstacle_radius) {
		push();
		FlipStepParams sp{dt, gravity, flip_ratio, num_pressure_iters, num_particle_iters, over_relaxation,
		                  compensate_drift ? 1 : 0, separate_particles ? 1 : 0, obstacle_x, obstacle_y, obstacle_vel_x,
		                  obstacle_vel_y, obstacle_radius, solver_order, update_colors ? 1 : 0};
		check(flip_step(sim_, &sp, 1), "flip_step");
	}

	// device-side FluidUI::setObstacle (main.cpp:95-126): avoids the host round trip of s,u,v
	void setObstacle(float x, float y, float vx, float vy, float radius) {
		push();
		check(flip_set_obstacle(sim_, x, y, vx, vy, radius), "flip_set_obstacle");
	}

	// refresh the mirrors the renderer reads (main.cpp:173-194)
	void pull_render_state(bool cells = true) {
		check(flip_readback(sim_, FLIP_PARTICLE_POS, particle_pos.data(), 2 * (size_t)num_particles), "readback particle_pos");
		if (update_colors) {
			check(flip_readback(sim_, FLIP_PARTICLE_COLOR, particle_color.data(), 3 * (size_t)num_particles), "readback particle_color");
			if (cells) check(flip_readback(sim_, FLIP_CELL_COLOR, cell_color.data(), cell_color.size()), "readback cell_color");
		}
	}
	void download(int field, void* host, size_t count) { check(flip_readback(sim_, field, host, count), "flip_readback"); }
	float particle_rest_density() {
		float r = 0;
		check(flip_get_scalar(sim_, FLIP_SCALAR_REST_DENSITY, &r), "flip_get_scalar");
		return r;
	}
	void synchronize() { check(flip_synchronize(sim_), "flip_synchronize"); }
	FlipSim* handle() { return sim_; }

private:
	FlipSim* sim_ = nullptr;
	bool dirty_particles_ = true, dirty_grid_ = true;
};

struct FluidGPU {
	float density = 0, h = 0, h1 = 0;
	std::vector<float> u, v, m, pressure;  // host mirrors, y-fastest (fluid.h:94-96)
	std::vector<unsigned char> solid;      // the reference's bool[]
	int solver_order = FLIP_SOLVER_LEX_WAVEFRONT;
	enum { U_FIELD = 0, V_FIELD, S_FIELD };

	FluidGPU(int x, int y, float d, float h_, int device = -1) {  // Fluid(x,y,d,h) fluid.h:41
		check(euler_create(x, y, d, h_, device, &sim_), "euler_create");
		EulerDims dm;
		check(euler_get_dims(sim_, &dm), "euler_get_dims");
		num_x_ = dm.num_x; num_y_ = dm.num_y; num_cells_ = dm.num_cells; density = dm.density; h = dm.h; h1 = dm.h1;
		u.assign(num_cells_, 0.f); v.assign(num_cells_, 0.f); m.assign(num_cells_, 1.f); pressure.assign(num_cells_, 0.f);
		solid.assign(num_cells_, 0);
	}
	FluidGPU(const FluidGPU&) = delete;
	FluidGPU& operator=(const FluidGPU&) = delete;
	~FluidGPU() { euler_destroy(sim_); }
	int getNumX() const { return num_x_; }
	int getNumY() const { return num_y_; }
	int getNumCells() const { return num_cells_; }
	int ix(int i, int j) const { return i * num_y_ + j; }
	void mark_dirty() { dirty_ = true; }
	void push() {
		if (!dirty_) return;
		check(euler_upload(sim_, EULER_U, u.data(), u.size()), "upload u");
		check(euler_upload(sim_, EULER_V, v.data(), v.size()), "upload v");
		check(euler_upload(sim_, EULER_M, m.data(), m.size()), "upload m");
		check(euler_upload(sim_, EULER_SOLID, solid.data(), solid.size()), "upload solid");
		dirty_ = false;
	}
	void solveIncompressibility(int num_iter, float dt) { push(); check(euler_project(sim_, num_iter, dt, solver_order), "euler_project"); }
	void extrapolate() { push(); check(euler_extrapolate(sim_), "euler_extrapolate"); }
	void advectVel(float dt) { push(); check(euler_advect_vel(sim_, dt), "euler_advect_vel"); }
	void advectSmoke(float dt) { push(); check(euler_advect_smoke(sim_, dt), "euler_advect_smoke"); }
	// device FluidUI::setObstacle(x, y, r, reset, dt), fluid_ui.h:24-62; the host mirrors of solid/m/u/v are NOT
	// refreshed (call pull_velocity()/pull_render_state() or download what you need)
	int obstacle_x = 0, obstacle_y = 0;
	void setObstacle(int x, int y, int r, bool reset = true, float dt = 1) {
		push();
		float vx = 0.f, vy = 0.f;
		if (!reset) { vx = (x - obstacle_x) / dt; vy = (y - obstacle_y) / dt; }
		obstacle_x = x; obstacle_y = y;
		check(euler_set_obstacle(sim_, x, y, r, vx, vy), "euler_set_obstacle");
	}
	void pull_render_state() {  // what fluid_ui.h:116-150 reads
		check(euler_readback(sim_, EULER_M, m.data(), m.size()), "readback m");
		check(euler_readback(sim_, EULER_PRESSURE, pressure.data(), pressure.size()), "readback pressure");
	}
	void pull_velocity() {
		check(euler_readback(sim_, EULER_U, u.data(), u.size()), "readback u");
		check(euler_readback(sim_, EULER_V, v.data(), v.size()), "readback v");
	}
	float sampleField(float x, float y, int field) {
		float out = 0;
		check(euler_sample_field(sim_, field, &x, &y, &out, 1), "euler_sample_field");
		return out;
	}
	void synchronize() { check(euler_synchronize(sim_), "euler_synchronize"); }
	EulerSim* handle() { return sim_; }

private:
	EulerSim* sim_ = nullptr;
	int num_x_ = 0, num_y_ = 0, num_cells_ = 0;
	bool dirty_ = true;
};

}  // namespace flipgpu
