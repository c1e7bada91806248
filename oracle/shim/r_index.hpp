// stand-in for the un-vendored third-party header of the same name (test infrastructure; see spumoni_ref_shim.hpp)
#pragma once
#include "spumoni_ref_shim.hpp"
