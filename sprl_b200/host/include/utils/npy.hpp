// utils/npy.hpp -- include-path compatibility with the reference's cpp/src/utils/npy.hpp (the writer is sprl_write_npy_f32): the
// declarations a worker or match main uses live in sprl/veneer.hpp (a handle layer over libsprl_b200.so).
#ifndef SPRL_B200_COMPAT_UTILS_NPY_HPP
#define SPRL_B200_COMPAT_UTILS_NPY_HPP
#include "../sprl/veneer.hpp"
#endif
