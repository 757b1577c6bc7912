// src/slam/cuda/ module header: the C ABI lives in the repository's include/ directory.
#include "../../../../include/mcl_cuda.h"
