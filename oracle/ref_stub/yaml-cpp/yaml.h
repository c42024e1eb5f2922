// yaml-cpp/yaml.h — STUB (oracle/_ref): the hot path never touches YAML; params.h only includes it.
#pragma once
