#pragma once
#include "ImathBox.h"
