// agents/HumanAgent.hpp -- include-path compatibility with the reference's cpp/src/agents/HumanAgent.hpp (interactive play: out of scope, nothing to declare): the
// declarations a worker or match main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_AGENTS_HUMANAGENT_HPP
#define SPRL_B200_COMPAT_AGENTS_HUMANAGENT_HPP
#include "../sprl/veneer.hpp"
#endif
