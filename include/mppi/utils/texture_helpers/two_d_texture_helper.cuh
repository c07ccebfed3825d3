// Source-compatibility forwarder: the reference's include path, served by the B200 host layer.
#pragma once
#include <mppi_b200/utils/texture_helpers/two_d_texture_helper.hpp>
