/* see glm.hpp in this directory (stand-in, oracle/_ref build only) */
#pragma once
#include <glm/glm.hpp>
