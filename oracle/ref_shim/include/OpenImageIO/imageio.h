#pragma once
#include "ustring.h"
