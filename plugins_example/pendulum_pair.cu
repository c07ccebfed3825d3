/*
 * plugins_example/pendulum_pair.cu — an OUT-OF-TREE (dynamics, cost) pair: a torque-driven pendulum with a quadratic cost.
 * Shows what a user of the reference does to bring their own model: write the device twins with the static methods of
 * mppi-generic_b200/csrc/plugins/dynamics.cuh / costs.cuh (the names and meaning of the reference's device methods:
 * computeDynamics / step / enforceConstraints, computeStateCost / terminalCost; per-thread register arrays instead of
 * shared-memory slices), give them POD parameter structs, and register the pair. Built by plugins_example/build.sh into
 * libmppi_plugin_pendulum.so; loaded with mppib_load_plugin(path); engines are then created with dynamics_id = cost_id = 1000.
 * Nothing in libmppi_b200.so is edited or rebuilt.
 */
#include "../mppi-generic_b200/csrc/engine_internal.cuh"

struct pendulum_dyn_params
{
  mppib_control_limits lim;  // every dynamics blob starts with the control limits (enforceConstraints)
  float mass, length, damping, gravity;
};
struct pendulum_cost_params
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM];  // CostParams<C> prefix (cost.cuh:18-30)
  float discount;
  float angle_coeff, rate_coeff, goal_angle, terminal_coeff;
};

// state (theta, theta_dot), control torque, output = state
struct PendulumDynamics : public mppib::plugins::Dynamics<PendulumDynamics, pendulum_dyn_params, 2, 1, 2>
{
  __device__ static __forceinline__ void computeDynamics(const Params& p, const float*, const float* x, const float* u,
                                                         float* xdot)
  {
    xdot[0] = x[1];
    xdot[1] = (u[0] - p.damping * x[1] - p.mass * p.gravity * p.length * __sinf(x[0])) / (p.mass * p.length * p.length);
  }
};
struct PendulumCost : public mppib::plugins::Cost<PendulumCost, pendulum_cost_params>
{
  __device__ static __forceinline__ float computeStateCost(const Params& p, const Aux&, const float*, const float* y, int, int*)
  {
    const float da = y[0] - p.goal_angle;
    return p.angle_coeff * da * da + p.rate_coeff * y[1] * y[1];
  }
  __device__ static __forceinline__ float terminalCost(const Params& p, const Aux& a, const float* y)
  {
    return p.terminal_coeff * computeStateCost(p, a, nullptr, y, 0, nullptr);
  }
};

extern "C" int mppib_plugin_init(void)
{
  return register_pair<PendulumDynamics, PendulumCost>(MPPIB_USER_ID_BASE + 0, MPPIB_USER_ID_BASE + 0);
}
