#pragma once
#include "../../../../ac_int_shim.hpp"
