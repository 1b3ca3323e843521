// Stand-in (nothing of boost::date_time is used on the tracking path).
#pragma once
