#pragma once
#include "../ImathColor.h"
