/* TEST INFRASTRUCTURE ONLY: see sync_block.h */
#include "sync_block.h"
