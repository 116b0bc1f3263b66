// boost/config.hpp — the one macro srtb uses from Boost.Config
#pragma once
#ifndef BOOST_FORCEINLINE
#define BOOST_FORCEINLINE inline __attribute__((always_inline))
#endif
