// placeholder, replaced by the tcgen05 kernels
#pragma once
#include "common.cuh"
namespace srhep { struct Bf16Weights {}; }
