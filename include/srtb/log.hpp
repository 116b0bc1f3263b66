// srtb/log.hpp — levelled stderr logger with the reference's macro names
// (reference: userspace/include/srtb/log/log.hpp:23-25,41-59,125-128: levels NONE..DEBUG,
// env SRTB_LOG_LEVEL, elapsed-seconds prefix). Re-hosted: one mutex-guarded line per statement.
#pragma once
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>

namespace srtb {
namespace log {

enum class levels : int { NONE = 0, ERROR = 1, WARNING = 2, INFO = 3, DEBUG = 4 };

inline levels initial_level() {
  if (const char* e = std::getenv("SRTB_LOG_LEVEL")) {
    const int v = std::atoi(e);
    if (v >= 0 && v <= 4) return static_cast<levels>(v);
  }
  return levels::INFO;
}
inline levels current_level = initial_level();
inline const auto start_time = std::chrono::steady_clock::now();
inline std::mutex sink_mutex;

class line {
 public:
  line(levels lv, const char* tag) : enabled_{static_cast<int>(lv) <= static_cast<int>(current_level)} {
    if (enabled_) {
      const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_time).count();
      buf_ << "[" << t << "] " << tag;
    }
  }
  ~line() {
    if (enabled_) {
      std::lock_guard<std::mutex> g{sink_mutex};
      std::cerr << buf_.str() << std::endl;
    }
  }
  template <typename T>
  line& operator<<(const T& v) {
    if (enabled_) buf_ << v;
    return *this;
  }

 private:
  bool enabled_;
  std::ostringstream buf_;
};

}  // namespace log
// the reference ends statements with `<< srtb::endl`; here the line flushes on destruction
struct endl_t {};
inline constexpr endl_t endl{};
inline log::line& operator<<(log::line& l, endl_t) { return l; }
inline log::line& operator<<(log::line&& l, endl_t) { return l; }
}  // namespace srtb

#define SRTB_LOGE ::srtb::log::line(::srtb::log::levels::ERROR, "[E]")
#define SRTB_LOGW ::srtb::log::line(::srtb::log::levels::WARNING, "[W]")
#define SRTB_LOGI ::srtb::log::line(::srtb::log::levels::INFO, "[I]")
#define SRTB_LOGD ::srtb::log::line(::srtb::log::levels::DEBUG, "[D]")
