// stub: include/Utils.hpp declares one function returning std::vector<boost::filesystem::path>
#pragma once
#include <string>
#include <vector>
namespace boost { namespace filesystem { class path { public: std::string s; }; } }
