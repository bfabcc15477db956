#include "b200_linux_config.h"
