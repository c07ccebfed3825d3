// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
#pragma once
#include <mppi_b200/dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.hpp>
