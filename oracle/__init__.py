"""oracle — TEST INFRASTRUCTURE ONLY. See oracle/loader.py."""
