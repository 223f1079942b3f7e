"""CPU oracle of the reference path — TEST INFRASTRUCTURE ONLY (see gpd_oracle.cpp)."""
