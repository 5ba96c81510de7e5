#pragma once
#include "../../../CL/sycl.hpp"
