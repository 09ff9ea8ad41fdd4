// stand-in for the reference header of the same name (tests only): see standin_types.h
#pragma once
#include "standin_types.h"
