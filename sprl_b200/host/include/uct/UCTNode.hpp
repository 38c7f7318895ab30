// uct/UCTNode.hpp -- include-path compatibility with the reference's cpp/src/uct/UCTNode.hpp: the
// declarations a worker or match main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_UCT_UCTNODE_HPP
#define SPRL_B200_COMPAT_UCT_UCTNODE_HPP
#include "../sprl/veneer.hpp"
#endif
