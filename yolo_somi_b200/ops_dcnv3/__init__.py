"""Host-side mirror of the reference package ``models/ops_dcnv3`` (functions/ + modules/)."""
