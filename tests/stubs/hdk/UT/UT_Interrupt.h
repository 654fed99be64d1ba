#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
class UT_AutoInterrupt {
   public:
	explicit UT_AutoInterrupt(const char* opname);
	~UT_AutoInterrupt();
	bool wasInterrupted(int percent = -1);
};
