#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <string>
class UT_String {
   public:
	UT_String();
	UT_String(const char* s);
	bool isstring() const;
	operator const char*() const;
	bool operator==(const char* s) const;
	bool multiMatch(const char* pattern, bool caseSensitive = true) const;
	std::string toStdString() const;
};
