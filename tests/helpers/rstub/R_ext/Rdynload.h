#include "../rstub.h"
