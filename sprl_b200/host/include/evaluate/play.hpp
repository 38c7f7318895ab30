// evaluate/play.hpp -- include-path compatibility with the reference's cpp/src/evaluate/play.hpp: the
// declarations a worker or match main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_EVALUATE_PLAY_HPP
#define SPRL_B200_COMPAT_EVALUATE_PLAY_HPP
#include "../sprl/veneer.hpp"
#include "../utils/Timer.hpp"     // as cpp/src/evaluate/play.hpp:8 does; Time.cpp relies on it
#endif
