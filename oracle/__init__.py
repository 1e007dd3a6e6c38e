"""Parity oracle package (TEST INFRASTRUCTURE ONLY -- never imported by cobaya_b200)."""
