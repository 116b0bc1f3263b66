// srtb/pipeline/framework/pipe.hpp (shim) — oracle/_ref calls the pipe functors directly and never
// starts threads, so only the standard headers the pipe headers rely on are provided
#pragma once
#include <optional>
#include <stop_token>
#include <thread>
#include "srtb/log/log.hpp"
