#pragma once
#include "tower.cuh"
