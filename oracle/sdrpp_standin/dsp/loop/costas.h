// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/loop/costas.h>.  The reference only needs it to pull in loop::PLL
// (PI4DQPSK_COSTAS derives from PLL, /root/reference/src/dsp/pi4dqpsk_costas.h:25).
#pragma once
#include "pll.h"
