/* stand-in for boost/format.hpp: just enough of boost::format / boost::str for the file names
 * mimo/framing.cc builds ("%s%d.dat" style, positional in order of operator%). */
#ifndef RUB_SHIM_BOOST_FORMAT_HPP
#define RUB_SHIM_BOOST_FORMAT_HPP
#include <sstream>
#include <string>
#include <vector>
namespace boost {
class format {
  std::string fmt_;
  std::vector<std::string> args_;
 public:
  explicit format(const char *f) : fmt_(f) {}
  explicit format(const std::string &f) : fmt_(f) {}
  template <class T> format &operator%(const T &v) { std::ostringstream o; o << v; args_.push_back(o.str()); return *this; }
  std::string str() const {
    std::string out;
    size_t a = 0;
    for (size_t i = 0; i < fmt_.size(); i++) {
      if (fmt_[i] != '%') { out.push_back(fmt_[i]); continue; }
      if (i + 1 < fmt_.size() && fmt_[i + 1] == '%') { out.push_back('%'); i++; continue; }
      size_t j = i + 1;  // skip flags / width / precision up to the conversion character
      while (j < fmt_.size() && !isalpha((unsigned char)fmt_[j])) j++;
      if (a < args_.size()) out += args_[a++];
      i = j;
    }
    return out;
  }
};
inline std::string str(const format &f) { return f.str(); }
inline std::ostream &operator<<(std::ostream &o, const format &f) { return o << f.str(); }
}  // namespace boost
#endif
