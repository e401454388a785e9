// Shim for <RBC/data_types.hpp> of nlamprian/RandomBallCover (not vendored by the reference).
// Layout stated by the reference at kernels/icp_kernels.cl:34-38 (dist_id).
#pragma once
#include <CL/cl.hpp>
typedef struct { cl_float dist; cl_uint id; } rbc_dist_id;
