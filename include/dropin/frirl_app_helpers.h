/* forwarding header: the whole reference API lives in friq_refapi.h */
#include "friq_refapi.h"
