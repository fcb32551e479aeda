"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference hot path (see oracle/sam2_path.py,
oracle/cc_oracle.c).  Never imported by the product package."""
