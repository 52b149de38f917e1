/* Stand-in for boost::thread_group (ChunkManagerHelper.h:212-225) over std::thread. Test infrastructure. */
#pragma once
#include <thread>
#include <vector>
#include <mutex>
namespace boost {
class thread_group {
  std::vector<std::thread> t_;
 public:
  template <typename F> void create_thread(F&& f) { t_.emplace_back(std::forward<F>(f)); }
  void join_all() { for (auto& t : t_) t.join(); t_.clear(); }
  ~thread_group() { join_all(); }
};
using mutex = std::mutex;
template <typename M> using lock_guard = std::lock_guard<M>;
}  // namespace boost
