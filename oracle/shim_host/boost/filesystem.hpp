// Host-side stand-in for <boost/filesystem.hpp>, enough for the unmodified reference APD.cpp. Test infrastructure only.
#ifndef APD_ORACLE_SHIM_HOST_BOOST_FS_HPP
#define APD_ORACLE_SHIM_HOST_BOOST_FS_HPP
#include <cstdlib>
#include <fstream>
#include <ostream>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
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
inline std::ostream &operator<<(std::ostream &o, const path &p) { return o << p.string(); }
class ifstream : public std::ifstream {
public:
	ifstream() {}
	explicit ifstream(const path &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p.string(), m) {}
};
class ofstream : public std::ofstream {
public:
	ofstream() {}
	explicit ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.string(), m) {}
};
inline bool exists(const path &p) { struct stat st; return stat(p.string().c_str(), &st) == 0; }
inline bool create_directory(const path &p) { return mkdir(p.string().c_str(), 0777) == 0; }
// APD_SHIM_KEEP_FILES=1 (tests only): main()'s clean-up of the intermediate .dmb / .bin files (main.cpp:219-226) becomes a no-op, so a
// test can read what the reference program wrote
inline bool remove(const path &p) { if (getenv("APD_SHIM_KEEP_FILES")) return true; return unlink(p.string().c_str()) == 0; }
}}  // namespace boost::filesystem
#endif
