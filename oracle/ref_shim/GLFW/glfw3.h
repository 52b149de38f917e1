/* Stand-in for GLFW: Runtimes/Helper/Timer.h only needs a monotonic clock. Test infrastructure. */
#pragma once
#include <chrono>
inline double glfwGetTime() {
  using namespace std::chrono;
  static const steady_clock::time_point t0 = steady_clock::now();
  return duration<double>(steady_clock::now() - t0).count();
}
