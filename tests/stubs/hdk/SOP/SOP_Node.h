#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <GU/GU_Detail.h>
#include <OP/OP_Operator.h>
#include <UT/UT_String.h>
enum SOP_ErrorCodes { SOP_MESSAGE = 0 };
class SOP_Node : public OP_Node {
   public:
	virtual const char* inputLabel(unsigned idx) const;

   protected:
	SOP_Node(OP_Network* net, const char* name, OP_Operator* op);
	virtual OP_ERROR cookMySop(OP_Context& context) = 0;
	OP_ERROR duplicateSource(unsigned index, OP_Context& context);
	exint evalInt(const char* parm, int vi, fpreal t) const;
	void evalString(UT_String& val, const char* parm, int vi, fpreal t) const;
	void setInt(const char* parm, int vi, fpreal t, exint value);
	void addError(int code, const char* msg = nullptr);
	void addMessage(int code, const char* msg = nullptr);
	OP_ERROR error();
	GU_Detail* gdp;
};
