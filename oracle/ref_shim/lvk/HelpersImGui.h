/* ImGui helpers are not on the path; Runtimes/Shape/Shape.h:13 includes this header unconditionally. Test infrastructure. */
#pragma once
#include <lvk/LVK.h>
