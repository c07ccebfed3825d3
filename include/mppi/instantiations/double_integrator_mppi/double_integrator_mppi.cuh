// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
#pragma once
#include <mppi_b200/controllers/MPPI/mppi_controller.hpp>
#include <mppi_b200/controllers/Tube-MPPI/tube_mppi_controller.hpp>
#include <mppi_b200/dynamics/double_integrator/di_dynamics.hpp>
#include <mppi_b200/cost_functions/double_integrator/double_integrator_circle_cost.hpp>
