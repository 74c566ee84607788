"""The reference's `lib/model` packages that sit on the region-level hot path, same names and signatures."""
