// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/clock_recovery/mm.h>.  Included by the reference headers but nothing
// from it is used on the hot path; it only has to exist and
// bring in what upstream's version transitively provides.
#pragma once
#include "../processor.h"
#include "../loop/phase_control_loop.h"
#include "../taps/windowed_sinc.h"
#include "../multirate/polyphase_bank.h"
#include "../math/step.h"
#include <volk/volk.h>
