// srtb/log/log.hpp (shim) — SRTB_LOG* swallow their operands
#pragma once
#include <string>
namespace srtb {
struct null_log {
  template <class T>
  null_log& operator<<(const T&) { return *this; }
};
struct endl_t {};
inline constexpr endl_t endl{};
namespace log {
enum class levels : int { NONE = 0, ERROR = 1, WARNING = 2, INFO = 3, DEBUG = 4 };
inline levels current_level = levels::NONE;
}  // namespace log
}  // namespace srtb
#define SRTB_LOGE ::srtb::null_log{}
#define SRTB_LOGW ::srtb::null_log{}
#define SRTB_LOGI ::srtb::null_log{}
#define SRTB_LOGD ::srtb::null_log{}
