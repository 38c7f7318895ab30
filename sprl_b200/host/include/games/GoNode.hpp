// games/GoNode.hpp -- include-path compatibility with the reference's cpp/src/games/GoNode.hpp: the
// declarations a worker main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_GAMES_GONODE_HPP
#define SPRL_B200_COMPAT_GAMES_GONODE_HPP
#include "../sprl/veneer.hpp"
#endif
