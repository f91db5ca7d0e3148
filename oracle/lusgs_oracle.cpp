// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// (LU-SGS restatement is added with the sweep kernel; see rho_oracle.cpp header.)
