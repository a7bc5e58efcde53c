#include "ahf_oracle.h"
