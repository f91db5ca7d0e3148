"""ORACLE / TEST INFRASTRUCTURE -- not product code.  See rho_oracle.cpp."""
