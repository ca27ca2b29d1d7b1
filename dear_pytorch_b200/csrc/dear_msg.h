// dear_msg.h — tiny string builder for error messages.
//
// std::ostringstream formats NUMBERS through locale facets, and in this image that crashes inside extensions
// (the extension is compiled against one libstdc++ and runs against the interpreter's; streaming an integer into
// an ostringstream segfaults — found with BucketSet::rs_plan on the first hardware run of round 2).  Messages are
// built with std::to_string instead, which has no locale dependency.
#pragma once
#include <string>
#include <type_traits>

namespace dear {

class Msg {
 public:
  Msg& operator<<(const char* s) { if (s) s_ += s; return *this; }
  Msg& operator<<(const std::string& s) { s_ += s; return *this; }
  Msg& operator<<(char c) { s_ += c; return *this; }
  Msg& operator<<(bool b) { s_ += b ? "true" : "false"; return *this; }
  template <typename T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  Msg& operator<<(T v) { s_ += std::to_string(v); return *this; }
  template <typename T, typename std::enable_if<std::is_enum<T>::value, int>::type = 0>
  Msg& operator<<(T v) { s_ += std::to_string(static_cast<long long>(v)); return *this; }
  const std::string& str() const { return s_; }

 private:
  std::string s_;
};

}  // namespace dear
