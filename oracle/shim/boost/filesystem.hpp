// Minimal stand-in for <boost/filesystem.hpp> (paths are never touched on the
// PatchMatch path; Problem merely carries two of them). Test infrastructure only.
#ifndef APD_ORACLE_SHIM_BOOST_FS_HPP
#define APD_ORACLE_SHIM_BOOST_FS_HPP
#include <string>
#include <fstream>
namespace boost { namespace filesystem {
class path {
public:
	path() {}
	path(const char *s) : s_(s) {}
	path(const std::string &s) : s_(s) {}
	const std::string &string() const { return s_; }
	path operator/(const path &o) const { return path(s_ + "/" + o.s_); }
private:
	std::string s_;
};
typedef std::ifstream ifstream;
typedef std::ofstream ofstream;
}}  // namespace boost::filesystem
#endif
