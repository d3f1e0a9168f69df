"""Mirror of the reference's ``experiments/base`` training loop on top of the device-resident agent."""
