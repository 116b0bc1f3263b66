// srtb/util/termination_handler.hpp (shim) — no Boost.Stacktrace here; nothing to install
#pragma once
