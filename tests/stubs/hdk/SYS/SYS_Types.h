// TEST INFRASTRUCTURE — declaration-only stand-ins for the handful of Houdini Development Kit classes that
// vqvdb_b200/cpp/sop/SOP_VQVDB_B200.cpp touches, so that the shim can be syntax-checked without the HDK
// (tests/test_sop_shim.py).  Signatures follow the HDK 20.5 headers of the same names; nothing here is implemented.
#pragma once
#include <cstdint>
using fpreal = double;
using exint = int64_t;
enum UT_ErrorSeverity { UT_ERROR_NONE = 0, UT_ERROR_MESSAGE, UT_ERROR_PROMPT, UT_ERROR_WARNING, UT_ERROR_ABORT, UT_ERROR_FATAL };
using OP_ERROR = UT_ErrorSeverity;
