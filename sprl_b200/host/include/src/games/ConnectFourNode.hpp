// The reference's unit test (cpp/tests/test_c4.cpp:1) includes "../src/games/ConnectFourNode.hpp" relative to cpp/tests;
// with -I<this tree>/include/games that path lands here.
#pragma once
#include "../../games/ConnectFourNode.hpp"
