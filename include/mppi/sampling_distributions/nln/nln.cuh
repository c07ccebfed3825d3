// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
#pragma once
#include <mppi_b200/sampling_distributions/nln/nln.hpp>
