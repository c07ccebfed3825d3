// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
#pragma once
#include <mppi_b200/cost_functions/quadrotor/quadrotor_quadratic_cost.hpp>
