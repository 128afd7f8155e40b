"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/pomp_oracle.c).  Never imported by the product."""
