// Stub of OUR OWN (see ../glm.hpp): glm::make_vec3 lives in glm.hpp of this stub.
#pragma once
#include "../glm.hpp"
