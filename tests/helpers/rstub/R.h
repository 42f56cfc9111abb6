#include "rstub.h"
