// Shared host-side helpers of libgenie_smem (error text, status codes).
#pragma once
#include <string>

namespace gsm {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
}  // namespace gsm
