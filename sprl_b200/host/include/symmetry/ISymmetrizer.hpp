// symmetry/ISymmetrizer.hpp -- include-path compatibility with the reference's cpp/src/symmetry/ISymmetrizer.hpp: the
// declarations a worker main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_SYMMETRY_ISYMMETRIZER_HPP
#define SPRL_B200_COMPAT_SYMMETRY_ISYMMETRIZER_HPP
#include "../sprl/veneer.hpp"
#endif
