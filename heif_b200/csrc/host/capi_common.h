// Shared by the C-ABI translation units: error capture so nothing unwinds across the boundary.
#pragma once
#include <cstdint>
#include <new>
#include <string>

#include "rbsp_reader.h"

namespace heic {

void set_last_error(const std::string& s);

template <class F>
int64_t guard(F&& fn) {
  try {
    return fn();
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return HEIC_E_NOMEM;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return HEIC_E_INVALID_ARG;
  } catch (...) {
    set_last_error("unknown error");
    return HEIC_E_INVALID_ARG;
  }
}

}  // namespace heic
