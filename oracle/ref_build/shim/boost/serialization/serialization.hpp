// Stand-in: the reference's serialize() member templates are never instantiated by oracle-R.
#pragma once
namespace boost { namespace serialization { class access {}; } }
