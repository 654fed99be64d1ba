#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <OP/OP_Operator.h>
class OP_AutoLockInputs {
   public:
	explicit OP_AutoLockInputs(OP_Node* node);
	~OP_AutoLockInputs();
	OP_ERROR lock(OP_Context& context);
};
