#pragma once
#include "ff_common.cuh"
