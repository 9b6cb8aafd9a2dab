// Shim of macr_b200/csrc/common.cuh for dev/tsan_sampler/run.sh (sampler.cu compiled as plain C++).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define MACR_OK 0
#define MACR_ERR_INVALID 1
namespace macr { inline int fail(int code, const char *fmt, ...) { return code; } }
#define MACR_CHECK_ARG(cond, ...) do { if (!(cond)) return ::macr::fail(MACR_ERR_INVALID, __VA_ARGS__); } while (0)
