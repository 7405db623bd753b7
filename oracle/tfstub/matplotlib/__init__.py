"""Empty matplotlib placeholder so the reference's `import matplotlib.pyplot as plt`
succeeds (plotting is out of scope).  TEST INFRASTRUCTURE ONLY."""
