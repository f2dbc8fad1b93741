// config/ConfigMap.h of the reference: class ConfigMap (defined in HydroRun.hpp, global namespace).
#include "../HydroRun.h"
