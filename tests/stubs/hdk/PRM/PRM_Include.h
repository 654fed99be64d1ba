#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <SYS/SYS_Types.h>
enum PRM_Type { PRM_STRING, PRM_FILE, PRM_INT, PRM_TOGGLE };
enum PRM_RangeFlag { PRM_RANGE_UI, PRM_RANGE_RESTRICTED };
class PRM_ChoiceList;
class PRM_Name {
   public:
	PRM_Name(const char* token = nullptr, const char* label = nullptr);
};
class PRM_Default {
   public:
	PRM_Default(fpreal f = 0, const char* s = nullptr);
};
class PRM_Range {
   public:
	PRM_Range(PRM_RangeFlag minflag, fpreal min, PRM_RangeFlag maxflag, fpreal max);
};
class PRM_Template {
   public:
	PRM_Template();
	PRM_Template(PRM_Type type, int vectorSize, PRM_Name* name, PRM_Default* defaults = nullptr, PRM_ChoiceList* choices = nullptr,
	             PRM_Range* range = nullptr);
};
extern PRM_Default PRMzeroDefaults[];
