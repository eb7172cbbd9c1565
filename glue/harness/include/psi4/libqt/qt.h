// HARNESS STAND-IN for psi4/libqt/qt.h:79-80: timers that remember how often each key was used.
#pragma once
#include <string>
namespace psi {
void timer_on(const std::string& key);
void timer_off(const std::string& key);
int harness_timer_count(const std::string& key);
}  // namespace psi
