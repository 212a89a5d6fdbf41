"""CPU oracle — TEST INFRASTRUCTURE ONLY.  See er_oracle.py.  The product never imports this package."""
