// Build shim for the verbatim reference build (oracle/_ref): the reference's
// games/GoNode.hpp:8 includes "../utils/zobrist.hpp" but the file on disk is
// utils/Zobrist.hpp (case-sensitive file systems fail).  Reached through
// -I<this dir>/x, so that "<x>/../utils/zobrist.hpp" resolves here.
#include "utils/Zobrist.hpp"
