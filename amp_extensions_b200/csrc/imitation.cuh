// placeholder until the imitation-reward kernel lands
#pragma once
namespace simstep { struct ImitConst { int dummy; }; }
