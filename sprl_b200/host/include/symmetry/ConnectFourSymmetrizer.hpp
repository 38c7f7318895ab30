// symmetry/ConnectFourSymmetrizer.hpp -- include-path compatibility with the reference's cpp/src/symmetry/ConnectFourSymmetrizer.hpp: the
// declarations a worker main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_SYMMETRY_CONNECTFOURSYMMETRIZER_HPP
#define SPRL_B200_COMPAT_SYMMETRY_CONNECTFOURSYMMETRIZER_HPP
#include "../sprl/veneer.hpp"
#endif
