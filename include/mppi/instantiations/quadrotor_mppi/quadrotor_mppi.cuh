// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
// (QuadrotorMapCost is not built: SURVEY §8 out-of-scope list / DESIGN.md §0.)
#pragma once
#include <mppi_b200/controllers/MPPI/mppi_controller.hpp>
#include <mppi_b200/dynamics/quadrotor/quadrotor_dynamics.hpp>
#include <mppi_b200/cost_functions/quadrotor/quadrotor_quadratic_cost.hpp>
