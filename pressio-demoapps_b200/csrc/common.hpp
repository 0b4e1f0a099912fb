// common.hpp -- error type shared by the host-side code; the C-ABI layer converts it to status codes.
#pragma once
#include <stdexcept>
#include <string>

namespace pda {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// mirror of the status enum in include/pda_b200.h (kept numeric here so host code does not need the C header)
enum : int { kOk = 0, kInvalid = 1, kIO = 2, kNoDevice = 3, kCuda = 4, kUnsupported = 5, kTooLarge = 6 };

}  // namespace pda
